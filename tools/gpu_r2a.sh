#!/bin/bash
# GPU call 1 of round 2: tests, north-star bench (C3) + C1, launch list and full captures on C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ free -g | head -2; nproc; nvidia-smi -L; } > gpurun_out/r2a_box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
timeout 300 python bench.py --workload plummer1m --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_c1.json 2> gpurun_out/r2a_bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2a_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_sph" -s 2 -c 2 -o gpurun_out/r2a_walk_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2a_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_keygen|k_sort_scatter|k_gather|k_links|k_upward" -c 12 -o gpurun_out/r2a_build_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2a_ncu3.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; head -c 600 gpurun_out/r2a_bench_c3.json
