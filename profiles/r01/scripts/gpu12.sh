cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python profiles/prof_run.py cfg3 f32 4000 100 > gpurun_out/prof_plain3.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:nwcwt_short2 -s 1 -c 1 -f -o gpurun_out/prof_cfg3_r01a python profiles/prof_run.py cfg3 f32 4000 100 > gpurun_out/ncu_cfg3.log 2>&1
tail -2 gpurun_out/ncu_cfg3.log
