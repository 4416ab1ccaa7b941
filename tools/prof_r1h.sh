set -x
B="python bench.py --steps 2 --warmup 1 --batch 64 --no-knn --no-cpu-baseline"
ncu --set full --import-source on --clock-control none -k regex:"k_grid_build|k_sp_" -s 9 -c 3 -o gpurun_out/prof_r1h_sp $B > gpurun_out/p1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_sl_" -s 6 -c 2 -o gpurun_out/prof_r1h_sl $B > gpurun_out/p2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_bow_" -s 6 -c 2 -o gpurun_out/prof_r1h_bow $B > gpurun_out/p3.log 2>&1
for n in sp sl bow; do ncu -i gpurun_out/prof_r1h_$n.ncu-rep --page raw --csv > gpurun_out/prof_r1h_${n}_raw.csv; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r1h.csv $B > gpurun_out/p4.log 2>&1
python bench.py > gpurun_out/bench_r1h_default.json 2> gpurun_out/bench_r1h_default.err
python bench.py --impl reference > gpurun_out/bench_r1h_reference.json 2> gpurun_out/bench_r1h_reference.err
tail -c 600 gpurun_out/bench_r1h_default.json; tail -c 400 gpurun_out/bench_r1h_reference.json
