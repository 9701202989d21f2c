#!/bin/bash
# GPU call 6 of round 2: multi-device handle (agb_multi_*), threaded AoS loops, gas-only SPH records — full GPU suite + benches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2f_box.txt
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2f_bench_c3.json 2> gpurun_out/r2f_bench_c3.err
tail -25 gpurun_out/r2f_pytest.log
