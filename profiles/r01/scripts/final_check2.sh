bash profiles/r01/scripts/final_check.sh
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --workload cfg5_26 --steps 3 --tuning 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg5_26 ms/step %.2f thr %s err %.1e' % (d['ms_per_step'], d['config']['threads'], d['parity_spot_check']['max_row_rel_l2']))"
timeout 300 python bench.py --workload cfg3_mean --steps 3 --tuning 2>&1 | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3_mean ms/step %.2f launches %d' % (d['ms_per_step'], d['gpu_launches']), d['parity_spot_check'])"
