cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 400 python bench.py --workload cfg3_mean --steps 5 --tuning > gpurun_out/bench_cfg3_mean.log 2>&1; echo rc=$?
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_cfg3_mean.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_cfg3_mean.log').read()[-800:])
else:
    d=json.loads(l[-1]); print('cfg3_mean: ms/step %.2f Gpts/s %.1f launches %d' % (d['ms_per_step'], d['value']/1e9, d['gpu_launches']), d['parity_spot_check'], {k:(round(v['ms_sum_of_launches'],1)) for k,v in d['roofline']['classes'].items()})
PY
timeout 400 python bench.py --workload cfg3 --steps 5 --tuning 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 ms/step %.2f' % d['ms_per_step'])"
NWCWT_NTHR_A=512 NWCWT_NTHR_B=512 timeout 400 python bench.py --workload cfg5_26 --steps 3 --tuning 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5_26 512thr ms/step %.2f' % d['ms_per_step'])"
timeout 400 python bench.py --workload cfg5_26 --steps 3 --tuning 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5_26 ms/step %.2f' % d['ms_per_step'])"
