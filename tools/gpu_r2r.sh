#!/bin/bash
# GPU call: extended-accuracy mode tests (+ quick check that the parity suite is untouched)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extended.py -m gpu -q -s > gpurun_out/r2r_pytest_ext.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2r_pytest_ext.log
tail -40 gpurun_out/r2r_pytest_ext.log


