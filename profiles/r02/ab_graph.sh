#!/bin/bash
# A/B of the CUDA-graph replay (NWCWT_GRAPH), stream count and launch-group size on the build with the batched pass-A gather.
set -u
O=gpurun_out/r02graph
mkdir -p $O
run() {   # name workload env...
  local name=$1 w=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --tuning > $O/${w}_$name.json 2> $O/${w}_$name.err
  python - <<E
import json
try:
    d = json.loads(open("$O/${w}_$name.json").read().strip().splitlines()[-1])
    print("$w $name: %.3f ms  launches/step %d  parity %s" % (d["ms_per_step"], d["gpu_launches"] // d["steps"], d["parity_spot_check"]))
except Exception as e:
    print("$w $name: failed", e, open("$O/${w}_$name.err").read()[-600:])
E
}
echo "== graph replay test"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "graph_replay or resampled_rows_at_cfg2 or long_rows_vs_golden" > $O/gputest_graph.log 2>&1; tail -3 $O/gputest_graph.log
for rep in 1 2; do
  run graph_$rep cfg2 NWCWT_GRAPH=1
  run nograph_$rep cfg2 NWCWT_GRAPH=0
done
run graph_s4 cfg2 NWCWT_GRAPH=1 NWCWT_STREAMS=4
run nograph_s4 cfg2 NWCWT_GRAPH=0 NWCWT_STREAMS=4
run graph_s2 cfg2 NWCWT_GRAPH=1 NWCWT_STREAMS=2
run graph_r24 cfg2 NWCWT_GRAPH=1 NWCWT_RING_MB=24
run graph_r32 cfg2 NWCWT_GRAPH=1 NWCWT_RING_MB=32
run graph_r24s4 cfg2 NWCWT_GRAPH=1 NWCWT_RING_MB=24 NWCWT_STREAMS=4
run graph_r16s4 cfg2 NWCWT_GRAPH=1 NWCWT_RING_MB=16 NWCWT_STREAMS=4
run graph cfg4 NWCWT_GRAPH=1
run nograph cfg4 NWCWT_GRAPH=0
run graph cfg5_20 NWCWT_GRAPH=1
run nograph cfg5_20 NWCWT_GRAPH=0
run graph cfg5_24 NWCWT_GRAPH=1
run nograph cfg5_24 NWCWT_GRAPH=0
