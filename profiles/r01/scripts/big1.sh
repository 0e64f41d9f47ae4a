cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
wl=$1; shift
env "$@" timeout 300 python bench.py --workload $wl --tuning --steps 5 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$wl $*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-400:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$wl $*", ': ms/step %.2f split %s radices %s thr %s' % (d['ms_per_step'], c['split'], c['radices'], c['threads']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run cfg4 NWCWT_STREAMS=2
run cfg4 NWCWT_BIG=1
run cfg4 NWCWT_BIG=1 NWCWT_CFG_B=4
run cfg4 NWCWT_BIG=1 NWCWT_STREAMS=1
run cfg4 NWCWT_STREAMS=1
run cfg2 NWCWT_STREAMS=2
run cfg2 NWCWT_BIG=1 NWCWT_SPLIT_N1=750
run cfg2 NWCWT_BIG=1 NWCWT_SPLIT_N1=750 NWCWT_CFG_B=4
run cfg2 NWCWT_BIG=3 NWCWT_SPLIT_N1=625
run cfg2 NWCWT_BIG=3 NWCWT_SPLIT_N1=625 NWCWT_CFG_B=4 NWCWT_CFG_A=4
run cfg2 NWCWT_BIG=1 NWCWT_SPLIT_N1=625
