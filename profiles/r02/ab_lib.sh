#!/bin/bash
# generic A/B of two library builds on the same box: ninwavelets_b200/libnwcwt_prev.so (previous) vs libnwcwt.so (new)
# (keep a copy of the current library as libnwcwt_prev.so before rebuilding; used for the stride-parameter experiment of
#  experiments_r02.md)
set -u
O=gpurun_out/r02lib
mkdir -p $O
for rep in 1 2 3; do
for v in prev new; do
  lib=$PWD/ninwavelets_b200/libnwcwt.so; [ $v = prev ] && lib=$PWD/ninwavelets_b200/libnwcwt_prev.so
  for w in cfg2 $( [ $rep = 1 ] && echo cfg4 cfg5_20 ); do
    NWCWT_LIB=$lib timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --tuning > $O/${w}_${v}_$rep.json 2> $O/${w}_${v}_$rep.err
    python - <<EOF
import json
try:
    d = json.loads(open("$O/${w}_${v}_$rep.json").read().strip().splitlines()[-1])
    c = d["roofline"]["classes"]
    print("$w $v rep $rep: %.3f ms  resample %.2f  parity %s" % (d["ms_per_step"], c.get("resample", {}).get("ms_sum_of_launches", 0), d["parity_spot_check"]["max_row_rel_l2"]))
except Exception as e:
    print("$w $v: failed", e)
EOF
  done
done
done
