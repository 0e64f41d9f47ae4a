# full GPU check of a build: parity tests, smoke, bench on cfg2 / cfg3 (run under gpurun from the repo root)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
for wl in cfg2 cfg3; do
timeout 600 python bench.py --workload $wl > gpurun_out/bench_$wl.log 2>&1; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$wl.log') if x.startswith('{')]
d=json.loads(l[-1]); print('$wl: ms/step %.2f Gpts/s %.1f frac %.4f e2e %.1f Gpts/s' % (d['ms_per_step'], d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9), d['parity_spot_check'], d['clocks'])
PY
done
