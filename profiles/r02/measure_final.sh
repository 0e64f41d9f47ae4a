#!/bin/bash
# r02 measurement pass on one B200 (run under gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/r02/measure_final.sh'
# Every ncu command runs only after the same program has exited 0 without ncu; numbers printed under ncu are never
# bench values.  Results land in gpurun_out/r02b/ (first pass of the round: gpurun_out/r02/) and are summarised into
# profiles/r02/ by profiles/r02/collect.py.
set -u
O=gpurun_out/r02b
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1

echo "== gpu tests"; ( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gputest.log 2>&1; echo "rc=$?" >> $O/gputest.log; tail -3 $O/gputest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log

echo "== bench cfg2 (default line)"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "rc=$?"
echo "== bench cfg3"
timeout 600 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "rc=$?"
echo "== bench cfg3_mean / cfg4 / cfg5 (tuning lines: device-resident only)"
for w in cfg3_mean cfg4 cfg5_16 cfg5_20 cfg5_22 cfg5_24 cfg5_26; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --tuning > $O/bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
echo "== bench cfg2 fp64 (tuning)"
timeout 300 python bench.py --dtype f64 --steps 5 --warmup 3 --tuning > $O/bench_cfg2_f64.json 2> $O/bench_cfg2_f64.err; echo "rc=$?"

echo "== ncu launch list of the bench command (durations + DRAM bytes per launch)"
# (NWCWT_GRAPH=0 under ncu: the same kernels launched from the host instead of replayed from the recorded graph; the list stops
# after the timed step = the 4th block of launches)
NWCWT_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
  -c 1320 --csv --log-file $O/launches_cfg2.csv python bench.py --steps 1 --warmup 3 --tuning > $O/ncu_launches.log 2>&1; echo "rc=$?"
gzip -f $O/launches_cfg2.csv

echo "== ncu --set full, cfg2 kernels (interpolation, decimated pass A / pass B)"
timeout 200 python profiles/prof_run.py cfg2 f32 16 100 > $O/prof_run_cfg2.log 2>&1; echo "rc=$?"
# launches of one transform: per launch group passA2 / passB2 / resample_dir; skip the first two transforms
NWCWT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nwcwt_resample -s 60 -c 3 -f -o $O/ncu_cfg2_rs \
  python profiles/prof_run.py cfg2 f32 16 100 > $O/ncu_cfg2_rs.log 2>&1; echo "rc=$?"
NWCWT_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nwcwt_pass.2_ -s 122 -c 6 -f -o $O/ncu_cfg2_ab \
  python profiles/prof_run.py cfg2 f32 16 100 > $O/ncu_cfg2_ab.log 2>&1; echo "rc=$?"
for r in ncu_cfg2_rs ncu_cfg2_ab; do
  [ -f $O/$r.ncu-rep ] && ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
ls -la $O
