cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size, getattr(p, "persisting_l2_cache_max_size", None), getattr(p, "access_policy_max_window_size", None))
PY
run() {
env "$@" timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
if not l: print("$*", 'FAILED', open('gpurun_out/bench_sweep.log').read()[-300:])
else:
    d=json.loads(l[-1]); c=d['config']; print("$*", ': ms/step %.2f rows/launch %d' % (d['ms_per_step'], c['rows_per_launch']), {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, '%.1e' % d['parity_spot_check']['max_row_rel_l2'])
PY
}
run NWCWT_STREAMS=2
run NWCWT_STREAMS=2 NWCWT_NO_L2_PERSIST=1
run NWCWT_STREAMS=2 NWCWT_RING_MB=32
run NWCWT_STREAMS=2 NWCWT_RING_MB=40
run NWCWT_STREAMS=1 NWCWT_RING_MB=64
run NWCWT_STREAMS=1 NWCWT_RING_MB=96
run NWCWT_STREAMS=1 NWCWT_RING_MB=96 NWCWT_NO_L2_PERSIST=1
run NWCWT_STREAMS=3 NWCWT_RING_MB=24
