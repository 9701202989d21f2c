#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "call_sequences" > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3a_pytest.log
tail -25 gpurun_out/r3a_pytest.log
