cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
wl=$1; shift
env "$@" timeout 300 python bench.py --workload $wl --tuning --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$wl $*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-400:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$wl $*", ': ms/step %.2f %.1f Gpts/s split %s radices %s thr %s' % (d['ms_per_step'], d['value']/1e9, c['split'], c['radices'], c['threads']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run cfg5_22 NWCWT_SPLIT_N1=4096
run cfg5_22 NWCWT_SPLIT_N1=8192
run cfg5_22 NWCWT_SPLIT_N1=1024
run cfg5_24 NWCWT_SPLIT_N1=8192
run cfg5_24 NWCWT_SPLIT_N1=2048
run cfg5_18 NWCWT_SPLIT_N1=256
run cfg5_18 NWCWT_SPLIT_N1=1024
run cfg5_20 NWCWT_SPLIT_N1=2048
run cfg5_20 NWCWT_SPLIT_N1=512
run cfg5_20 NWCWT_SPLIT_N1=4096
