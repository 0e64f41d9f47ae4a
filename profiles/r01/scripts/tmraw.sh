cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$*", ': ms/step %.2f' % (d['ms_per_step']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run NWCWT_STREAMS=1 NWCWT_RING_MB=200
run NWCWT_STREAMS=1 NWCWT_RING_MB=200 NWCWT_LIB=$GRAFT_REPO_ROOT/ninwavelets_b200/libnwcwt_raw.so
run NWCWT_STREAMS=2
run NWCWT_STREAMS=2 NWCWT_LIB=$GRAFT_REPO_ROOT/ninwavelets_b200/libnwcwt_raw.so
