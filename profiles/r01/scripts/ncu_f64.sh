cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python profiles/prof_run.py cfg2 f64 2 100 > gpurun_out/prof_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/prof_plain.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:nwcwt_pass.2 -s 8 -c 2 -f -o gpurun_out/prof_cfg2_f64_r01 python profiles/prof_run.py cfg2 f64 2 100 > gpurun_out/ncu_cfg2_f64.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_cfg2_f64.log
