#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py tests/test_gpu_extended.py -m gpu -q -k "multi or several or two_gpu" > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3e_pytest.log
tail -4 gpurun_out/r3e_pytest.log
