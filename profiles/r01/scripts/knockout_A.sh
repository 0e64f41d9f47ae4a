cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
env NWCWT_STREAMS=1 NWCWT_RING_MB=200 "$@" timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$*", ': ms/step %.2f' % (d['ms_per_step']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')})
PY
}
run A=1
for k in 3 4 5; do run NWCWT_LIB=$GRAFT_REPO_ROOT/ninwavelets_b200/libnwcwt_ko$k.so; done
