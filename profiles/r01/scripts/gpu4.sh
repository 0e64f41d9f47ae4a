cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for cfg in "1 48" "1 96" "1 200" "2 48" "2 64" "2 96" "3 48" "4 32" "2 200"; do
set -- $cfg
NWCWT_STREAMS=$1 NWCWT_RING_MB=$2 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_sweep.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sweep.log') if x.startswith('{')]
d=json.loads(l[-1]); print('streams $1 ringMB $2: value %.1f Gpts/s  ms/step %.2f rows/launch %d' % (d['value']/1e9, d['ms_per_step'], d['config']['rows_per_launch']), {k:(round(v['ms_sum_of_launches'],1), v['launches']) for k,v in d['roofline']['classes'].items() if k.startswith('inv')}, d['parity_spot_check']['max_row_rel_l2'])
PY
done
