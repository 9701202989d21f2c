#!/bin/bash
# GPU call 5 of round 2: gas-only SPH candidate records; k_sph at 4 CTAs/SM (variant E); device-resident loop probe on C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or against_oracle or late_upload or counter_mode or C3_gas16m-mixed or dudt or slices" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_E.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2e_bench_c3_E.json 2> gpurun_out/r2e_bench_c3_E.err
timeout 300 python tools/gpu_resident.py gas16m fused 16 > gpurun_out/r2e_resident.log 2>&1
tail -3 gpurun_out/r2e_pytest.log; tail -12 gpurun_out/r2e_resident.log
