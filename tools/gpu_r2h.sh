#!/bin/bash
# GPU call: three-word keys (63 levels) — deep-tree tests, full suite, device-resident loop probe on C3 past the blow-up
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coincident or closer_than or blown_up or tight or golden" > gpurun_out/r2h_pytest_deep.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2h_pytest_deep.log
tail -30 gpurun_out/r2h_pytest_deep.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2h_pytest.log
timeout 400 python tools/gpu_resident.py gas16m fused 24 > gpurun_out/r2h_resident.log 2>&1
tail -5 gpurun_out/r2h_pytest.log; tail -18 gpurun_out/r2h_resident.log
