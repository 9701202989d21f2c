#!/bin/bash
# GPU call 3 of round 2: AoS hand-over debug, full GPU tests on the new build (one-sweep sort, packed gather, level-synchronous upward pass, lane-per-node MAC), benches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/debug_aos.py galic22k > gpurun_out/r2c_debug.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=10 -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_c3.json 2> gpurun_out/r2c_bench_c3.err
timeout 200 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2c_bench_c1.json 2> gpurun_out/r2c_bench_c1.err
cat gpurun_out/r2c_debug.log; tail -15 gpurun_out/r2c_pytest.log
