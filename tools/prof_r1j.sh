# final profile pass of round 1 (run under gpurun, one GPU): launch list of the whole bench command, --set full captures of the
# kernels added in this session, default bench line + reference arm
set -x
B="python bench.py --steps 2 --warmup 1 --batch 64 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1j.csv $B > gpurun_out/p1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_fisheye_triangulate|k_fisheye_knn2" -c 4 -o gpurun_out/prof_r1j_fisheye \
    python bench.py --steps 2 --warmup 1 --batch 64 --workload tumvi --no-knn --no-cpu-baseline > gpurun_out/p2.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_knn2_merge_push|k_knn2_merge_wait|k_knn2_scan" -s 3 -c 3 -o gpurun_out/prof_r1j_knn_exchange \
    python bench.py --steps 1 --warmup 1 --batch 16 --no-match --no-cpu-baseline > gpurun_out/p3.log 2>&1
for n in fisheye knn_exchange; do ncu -i gpurun_out/prof_r1j_$n.ncu-rep --page raw --csv > gpurun_out/prof_r1j_${n}_raw.csv; done
python bench.py > gpurun_out/bench_r1j_default.json 2> gpurun_out/bench_r1j_default.err
python bench.py --impl reference > gpurun_out/bench_r1j_reference.json 2> gpurun_out/bench_r1j_reference.err
python bench.py --workload tumvi > gpurun_out/bench_r1j_tumvi.json 2> gpurun_out/bench_r1j_tumvi.err
python bench.py --workload tumvi --impl reference > gpurun_out/bench_r1j_tumvi_reference.json 2> gpurun_out/bench_r1j_tumvi_reference.err
tail -c 600 gpurun_out/bench_r1j_default.json; tail -c 400 gpurun_out/bench_r1j_reference.json
