#!/bin/bash
# environment-switch sweep on the shipped build (cfg2, device-resident): launch-group size above the default, interpolation grid
set -u
O=gpurun_out/r02env
mkdir -p $O
run() {
  local name=$1; shift
  env "$@" timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 3 --tuning > $O/cfg2_$name.json 2> $O/cfg2_$name.err
  python - <<EOF
import json
try:
    d = json.loads(open("$O/cfg2_$name.json").read().strip().splitlines()[-1])
    print("cfg2 $name: %.3f ms  launches/step %d  parity %s" % (d["ms_per_step"], d["gpu_launches"] // d["steps"], d["parity_spot_check"]["max_row_rel_l2"]))
except Exception as e:
    print("cfg2 $name: failed", e, open("$O/cfg2_$name.err").read()[-400:])
EOF
}
run default_1 NWCWT_GRAPH=1
run ring64 NWCWT_RING_MB=64
run ring96 NWCWT_RING_MB=96
run ring128 NWCWT_RING_MB=128
run ring96s4 NWCWT_RING_MB=96 NWCWT_STREAMS=4
run rsctas3 NWCWT_RS_CTAS=3
run default_2 NWCWT_GRAPH=1
