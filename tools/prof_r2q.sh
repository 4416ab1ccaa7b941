# profile pass of round 2 (run under gpurun, one GPU): launch list of one bench command, --set full capture of one steady-state step
# (both extractions + stereo matcher at 64 pairs per launch), raw page + per-kernel summary; SASS evidence of the TMA kernels
set -x
B="python bench.py --steps 2 --warmup 1 --batch 64 --no-knn --no-match --no-cpu-baseline --no-workloads"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2q.csv $B > gpurun_out/p1p.log 2>&1
ncu --set full --import-source on --clock-control none --launch-skip 81 --launch-count 27 -o gpurun_out/prof_r2q_pipeline $B > gpurun_out/p2p.log 2>&1
ncu -i gpurun_out/prof_r2q_pipeline.ncu-rep --page raw --csv > gpurun_out/ncu_full_r2q_pipeline_raw.csv
python tools/ncu_summary.py gpurun_out/ncu_full_r2q_pipeline_raw.csv 128 gpurun_out/traffic_r2q.json gpurun_out/launches_r2q_step.csv > gpurun_out/ncu_summary_r2q.md
cat gpurun_out/ncu_summary_r2q.md
cuobjdump -sass morb_slam_b200/lib/liborb_b200.so | grep -E "Function :|UTMALDG|UBLKCP|SYNCS" | grep -B1 -E "UTMALDG|UBLKCP" | grep -E "Function|UTMALDG" > gpurun_out/sass_tma_r2q.txt
head -20 gpurun_out/sass_tma_r2q.txt
