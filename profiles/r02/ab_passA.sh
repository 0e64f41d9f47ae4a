#!/bin/bash
# A/B of pass-A variants on cfg2 (and cfg4, cfg5_20): library builds side by side, selected with NWCWT_LIB.
#   base: the build measured in bench_lines_r02.jsonl; (default): four-slot gather with the loads issued first;
#   r96:  the same + 96-register cap for the 64-thread launch shape (10 instead of 8 resident CTAs per SM)
# (historical: libnwcwt_base.so = the build of commit 762d80c, libnwcwt_r96.so = k_long2_f32_c3.cu compiled with
#  -DNW_CFG3_MAXREG=96 and linked with the other objects; both were removed once the variant became the default)
set -u
O=gpurun_out/r02ab
mkdir -p $O
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
for rep in 1 2; do
for v in base "" r96; do
  lib=ninwavelets_b200/libnwcwt${v:+_$v}.so
  [ -f $lib ] || continue
  for w in cfg2 $( [ $rep = 1 ] && echo cfg4 cfg5_20 ); do
    NWCWT_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --tuning > $O/${w}_${v:-new}_$rep.json 2> $O/${w}_${v:-new}_$rep.err
    python - <<E
import json
try:
    d = json.loads(open("$O/${w}_${v:-new}_$rep.json").read().strip().splitlines()[-1])
    print("$w ${v:-new} rep $rep: %.3f ms  parity %s" % (d["ms_per_step"], d["parity_spot_check"]))
except Exception as e:
    print("$w ${v:-new}: failed", e)
E
  done
done
done
echo "== gpu parity tests on the new default build (long-row cases)"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resampled_rows or long_rows or sweep or awkward or 2_20 or properties" > $O/gputest_long.log 2>&1; tail -2 $O/gputest_long.log
