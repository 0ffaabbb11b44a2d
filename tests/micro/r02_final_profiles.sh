set -x
VSOM_TC_TIER=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc_kernel -s 1 -c 1 -o gpurun_out/r02_k2_pair_t1 -f python tests/profile_k2.py 524288 random > gpurun_out/r02_k2_pair_t1.log 2>&1
VSOM_TC_TIER=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:score_tc_kernel -s 1 -c 1 -o gpurun_out/r02_k2_pair_t2 -f python tests/profile_k2.py 524288 trained > gpurun_out/r02_k2_pair_t2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:batch_update -s 1 -c 1 -o gpurun_out/r02_k6_update -f python tests/bench_k6_quick.py > gpurun_out/r02_k6_update.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:local_bmu_rows -s 0 -c 1 -o gpurun_out/r02_k6_walk -f python tests/bench_k6_quick.py > gpurun_out/r02_k6_walk.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_pre_ncu.json 2> gpurun_out/r02_bench_pre_ncu.err && timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_under_ncu.log 2>&1
echo finished
