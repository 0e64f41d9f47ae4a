cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
run() {
wl=$1; dt=$2; shift; shift
env "$@" timeout 300 python bench.py --workload $wl --dtype $dt --tuning --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$wl $dt $*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-400:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$wl $dt $*", ': ms/step %.2f %.1f Gpts/s split %s thr %s rows/launch %s' % (d['ms_per_step'], d['value']/1e9, c['split'], c['threads'], c['rows_per_launch']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run cfg2 f64 A=1
run cfg2 f64 NWCWT_NO_NARROW=1
run cfg5_16 f32 A=1
run cfg5_16 f32 NWCWT_NO_NARROW=1
run cfg5_18 f32 A=1
run cfg5_20 f32 A=1
run cfg5_20 f32 NWCWT_NO_NARROW=1
run cfg5_22 f32 A=1
run cfg5_24 f32 A=1
run cfg4 f64 A=1
run cfg2 f32 A=1
