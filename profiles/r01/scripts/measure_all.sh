cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
tag=$1; shift
timeout 600 python bench.py "$@" > gpurun_out/bench_$tag.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$tag.log') if x.startswith('{')]
if not l: print("$tag", 'FAILED', open('gpurun_out/bench_$tag.log').read()[-600:])
else:
    d=json.loads(l[-1]); c=d['config']; e=d.get('e2e') or {}
    print("$tag", ': ms/step %.2f  %.1f Gpts/s  frac %.3f  path %s split %s e2e %s  parity %.1e cpu %s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], c['path'], c['split'], ('%.1f' % (e['value']/1e9)) if e.get('value') else None, d['parity_spot_check']['max_row_rel_l2'], (d.get('cpu_baseline') or {}).get('sample')))
PY
}
run cfg2_f64 --workload cfg2 --dtype f64 --steps 5
run cfg3_f64 --workload cfg3 --dtype f64 --steps 5 --tuning
run cfg4_f64 --workload cfg4 --dtype f64 --steps 5 --tuning
for k in 16 18 20 22 24 26; do run cfg5_$k --workload cfg5_$k --steps 3 --tuning; done
run cfg5_24_f64 --workload cfg5_24 --dtype f64 --steps 3 --tuning
run cfg2 --steps 20
