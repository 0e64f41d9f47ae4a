#!/bin/bash
# last call of the round: the full GPU suite and smoke() on the final build (clean rebuild of the committed sources), one bench line
set -u
O=gpurun_out/r02d
mkdir -p $O
( time timeout 400 python -m pytest tests -m gpu -x -q ) > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log; tail -6 $O/gputest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 100 python bench.py --steps 20 --warmup 3 --tuning > $O/bench_cfg2_tuning.json 2> $O/bench_cfg2_tuning.err
python -c "
import json
d = json.loads(open('$O/bench_cfg2_tuning.json').read().strip().splitlines()[-1])
print('cfg2 %.3f ms  %.1f G points/s  parity %s' % (d['ms_per_step'], d['value'] / 1e9, d['parity_spot_check']))"
