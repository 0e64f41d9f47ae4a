cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
env NWCWT_STREAMS=1 NWCWT_RING_MB=200 "$@" timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$*", ': ms/step %.2f thr %s batch %d' % (d['ms_per_step'], c['threads'], c['batch']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run NWCWT_NO_STATIC=1
run NWCWT_NO_STATIC=1 NWCWT_TPSH_B=0 NWCWT_NTHR_B=64 NWCWT_CFG_B=3
run NWCWT_NO_STATIC=1 NWCWT_TPSH_B=0 NWCWT_NTHR_B=96 NWCWT_CFG_B=2
run NWCWT_NO_STATIC=1 NWCWT_TPSH_B=0 NWCWT_NTHR_B=32 NWCWT_CFG_B=3
run NWCWT_NO_STATIC=1 NWCWT_TPSH_B=1 NWCWT_NTHR_B=64 NWCWT_CFG_B=3
run NWCWT_NO_STATIC=1 NWCWT_TPSH_A=0 NWCWT_NTHR_A=32 NWCWT_CFG_A=3
run NWCWT_NO_STATIC=1 NWCWT_TPSH_A=0 NWCWT_NTHR_A=64 NWCWT_CFG_A=3
