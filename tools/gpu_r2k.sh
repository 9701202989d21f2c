#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2k_bench_c3.json 2> gpurun_out/r2k_bench_c3.err
tail -12 gpurun_out/r2k_pytest.log
