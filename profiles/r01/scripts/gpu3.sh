cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
for ns in 2 1 3; do
NWCWT_STREAMS=$ns timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2_s$ns.log 2>&1; echo "bench streams=$ns rc=$?"; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_cfg2_s$ns.log') if x.startswith('{')]
d=json.loads(l[-1]); print('value %.1f Gpts/s  ms/step %.2f  frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'])); print({k:(round(v['ms_sum_of_launches'],2), v['launches']) for k,v in d['roofline']['classes'].items()}); print(d['parity_spot_check'], d['e2e']['value']/1e9)
PY
done
