#!/bin/bash
# the default bench line (and the reference arm) of the final build, after the graph-cache fix; plus the regression tests
set -u
O=gpurun_out/r02c
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plan_close or graph_replay or quirks or baseline_modes" > $O/gputest_fix.log 2>&1; tail -2 $O/gputest_fix.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench rc=$?"; tail -c 300 $O/bench_cfg2.err
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
python - <<EOF
import json
for f in ("bench_cfg2", "bench_cfg3", "bench_reference"):
    try:
        d = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "%.3f ms" % d["ms_per_step"], "%.2f G/s" % (d["value"] / 1e9), "e2e", (d.get("e2e") or {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "failed", e)
EOF
