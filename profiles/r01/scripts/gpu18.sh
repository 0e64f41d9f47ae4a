cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1 --warmup 3 > gpurun_out/bench_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_cfg2_r01.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_l.log 2>&1
echo "ncu rc=$?"; tail -c 300 gpurun_out/ncu_l.log; wc -l gpurun_out/launches_bench_cfg2_r01.csv
timeout 200 python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/prof_plain.log 2>&1 && \
NWCWT_STREAMS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_pass.2 -s 28 -c 4 -f -o gpurun_out/prof_cfg2_r01f python profiles/prof_run.py cfg2 f32 2 100 > gpurun_out/ncu_cfg2.log 2>&1
tail -1 gpurun_out/ncu_cfg2.log
