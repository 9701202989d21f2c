#!/bin/bash
# GPU call 4 of round 2: late-upload force path (tests + e2e), walk variant D (3 CTAs/SM), launch list and full captures of the build / k_sph on C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2d_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_bench_c3.err
AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_D.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2d_bench_c3_D.json 2> gpurun_out/r2d_bench_c3_D.err
AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_D.so timeout 200 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2d_bench_c1_D.json 2> gpurun_out/r2d_bench_c1_D.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2d_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2d_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sph|k_gather|k_keygen|k_upward_levels|k_links|k_sort_onesweep|k_pack_gas|k_gas_sum|k_gas_mark" -c 14 -o gpurun_out/r2d_build_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2d_ncu2.log 2>&1
tail -5 gpurun_out/r2d_pytest.log; head -c 1200 gpurun_out/r2d_bench_c3.json
