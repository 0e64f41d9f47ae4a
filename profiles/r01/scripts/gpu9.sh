cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); print("$*", ': ms/step %.2f thr %s batch %d' % (d['ms_per_step'], d['config']['threads'], d['config']['batch']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run NWCWT_STREAMS=1 NWCWT_RING_MB=200
run NWCWT_STREAMS=1 NWCWT_RING_MB=200 NWCWT_NO_PRUNE=1
timeout 200 python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/prof_plain.log 2>&1 && \
NWCWT_STREAMS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_passA2p -s 14 -c 2 -f -o gpurun_out/prof_cfg2_r01e python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log
