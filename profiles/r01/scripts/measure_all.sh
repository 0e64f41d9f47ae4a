# measurement table of a build: every config in fp32 / fp64 (tuning runs: no host leg, no CPU baseline)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
tag=$1; shift
timeout 600 python bench.py "$@" > gpurun_out/bench_$tag.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$tag.log') if x.startswith('{')]
if not l: print("$tag", 'FAILED', open('gpurun_out/bench_$tag.log').read()[-600:])
else:
    d=json.loads(l[-1]); c=d['config']
    print("| $tag | %s | %s | %s | %.2f | %.1f | %.3f | %.1e |" % (c['split'], c['radices'], c['threads'], d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['parity_spot_check']['max_row_rel_l2']))
PY
}
for dt in f32 f64; do
for wl in cfg2 cfg3 cfg4; do run ${wl}_$dt --workload $wl --dtype $dt --steps 5 --tuning; done
for k in 16 18 20 22 24 26; do run cfg5_${k}_$dt --workload cfg5_$k --dtype $dt --steps 3 --tuning; done
done
