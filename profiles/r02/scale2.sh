#!/bin/bash
# two GPUs of one box (gpurun --gpus 2): the weak-scaling bench command the driver uses, and ONE job sharded over the ranks by
# ninwavelets_b200.sharding.distributed_transform (strong scaling), cfg2 and the cfg5 2^20 sweep size
set -u
O=gpurun_out/r02scale
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
run() {
  local name=$1; shift
  timeout 400 "$@" > $O/$name.json 2> $O/$name.err
  python - <<EOF
import json
try:
    d = json.loads([l for l in open("$O/$name.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("$name: n_gpus %d  %s  %.3f ms/step  %.1f G points/s  parity %s" % (d["n_gpus"], d["scaling"], d["ms_per_step"], d["value"] / 1e9, d["parity_spot_check"]))
except Exception as e:
    print("$name: failed", e, open("$O/$name.err").read()[-800:])
EOF
}
if [ "${ONLY:-all}" != strong ]; then
run bench_1gpu_cfg2 python bench.py --gpus 1 --steps 10 --warmup 3 --tuning
run bench_2gpu_weak_cfg2 $TR bench.py --gpus 2 --steps 10 --warmup 3 --tuning
fi
run bench_2gpu_strong_cfg2 $TR bench.py --gpus 2 --steps 10 --warmup 3 --tuning --scaling strong
run bench_1gpu_strong_cfg5_20 python bench.py --gpus 1 --steps 3 --warmup 3 --tuning --scaling strong --workload cfg5_20
run bench_2gpu_strong_cfg5_20 $TR bench.py --gpus 2 --steps 3 --warmup 3 --tuning --scaling strong --workload cfg5_20
