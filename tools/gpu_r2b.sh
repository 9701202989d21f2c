#!/bin/bash
# GPU call 2 of round 2: debug of the galic22k failure, full test log on the base library, variant libraries A/B/C
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/debug_galic.py galic22k memcheck > gpurun_out/r2b_debug.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=10 -s > gpurun_out/r2b_pytest_base.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest_base.log
for v in A B C; do
  export AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_$v.so
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "golden or against_oracle or slices or tight or inactive or force_path or C1_plummer1m or C2_disk4m or tiny or lattice or counter_mode" > gpurun_out/r2b_pytest_$v.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2b_pytest_$v.log
  timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2b_bench_c3_$v.json 2> gpurun_out/r2b_bench_c3_$v.err
  timeout 200 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2b_bench_c1_$v.json 2> gpurun_out/r2b_bench_c1_$v.err
done
unset AGB200_LIB
tail -5 gpurun_out/r2b_debug.log; for v in base A B C; do tail -2 gpurun_out/r2b_pytest_$v.log; done
