cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gputests.log
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_cfg2.log
timeout 200 python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/prof_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_pass.2 -s 48 -c 4 -f -o gpurun_out/prof_cfg2_r01b python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_cfg2.log 2>&1
tail -3 gpurun_out/ncu_cfg2.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg2_r01b.csv python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log
