cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
run() {
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --workload cfg3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-400:])
else:
    d=json.loads(l[-1]); print("$*", ': ms/step %.2f Gpts/s %.1f frac %.3f path %s thr %s batch %d' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['config']['path'], d['config']['threads'], d['config']['batch']), '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run A=1
run NWCWT_TPSH_S=0
run NWCWT_NTHR_S=192
cp gpurun_out/bench_sweep.log gpurun_out/bench_cfg3.log
timeout 200 python profiles/prof_run.py cfg3 f32 4000 100 > gpurun_out/prof_plain3.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_short2 -s 1 -c 1 -f -o gpurun_out/prof_cfg3_r01b python profiles/prof_run.py cfg3 f32 4000 100 > gpurun_out/ncu_cfg3.log 2>&1
tail -1 gpurun_out/ncu_cfg3.log
