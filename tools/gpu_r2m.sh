#!/bin/bash
# GPU call: packed prefix scan in the walk, early gas-flag load, upward occupancy — parity subset + benches
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or against_oracle or late_upload or counter_mode or C3_gas16m-mixed or C1_plummer1m-mixed or dudt or slices or lattice or small_opening or extreme" > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2m_bench_c3.json 2> gpurun_out/r2m_bench_c3.err
timeout 300 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2m_bench_c1.json 2> gpurun_out/r2m_bench_c1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sph" -c 1 -o gpurun_out/r2m_sph_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2m_ncu2.log 2>&1
tail -3 gpurun_out/r2m_pytest.log
