cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -16 gpurun_out/gputests.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_cfg2.log') if x.startswith('{')]
d=json.loads(l[-1]); print('cfg2: ms/step %.2f Gpts/s %.1f frac %.4f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']), {k:(round(v['ms_sum_of_launches'],2), v['launches']) for k,v in d['roofline']['classes'].items()}, d['parity_spot_check'])
PY
timeout 300 python bench.py --steps 3 --warmup 3 --workload cfg4 > gpurun_out/bench_cfg4.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_cfg4.log') if x.startswith('{')]
d=json.loads(l[-1]); print('cfg4: ms/step %.2f Gpts/s %.1f frac %.4f' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac']), {k:(round(v['ms_sum_of_launches'],2), v['launches']) for k,v in d['roofline']['classes'].items()}, d['parity_spot_check'], d['config']['threads'], d['config']['radices'])
PY
