cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_cfg2.log') if x.startswith('{')]
d=json.loads(l[-1]); print('value %.1f Gpts/s  ms/step %.2f  frac %.4f' % (d['value']/1e9, d['ms_per_step'], d['roofline']['frac'])); print(d['roofline']['classes']); print(d['parity_spot_check'], d['e2e']['value']/1e9)
PY
timeout 200 python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/prof_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_pass.2 -s 48 -c 4 -f -o gpurun_out/prof_cfg2_r01c python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_cfg2.log 2>&1
tail -2 gpurun_out/ncu_cfg2.log
