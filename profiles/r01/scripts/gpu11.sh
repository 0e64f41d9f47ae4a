cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gputests.log
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
run NWCWT_NO_SHORT2=1
run NWCWT_TPSH_S=0
run NWCWT_TPSH_S=2
run NWCWT_NTHR_S=160
run NWCWT_NO_STATIC=1
cp gpurun_out/bench_sweep.log gpurun_out/bench_cfg3.log
