# historical: the persistent pass-B kernel / NW_TIMING hooks this script measured were removed again (see shape_sweep.md)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
run() {
env "$@" timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$*", ': ms/step %.2f rows/launch %d' % (d['ms_per_step'], c['rows_per_launch']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run NWCWT_STREAMS=1 NWCWT_RING_MB=200
run NWCWT_STREAMS=1 NWCWT_RING_MB=200 NWCWT_PERSIST_B=0
run NWCWT_STREAMS=2
run NWCWT_STREAMS=2 NWCWT_PERSIST_B=0
run NWCWT_STREAMS=2 NWCWT_RING_MB=100
run NWCWT_STREAMS=2 NWCWT_RING_MB=200
