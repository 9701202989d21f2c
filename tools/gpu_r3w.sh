#!/bin/bash
# last check inside the remaining budget: the driver's multi-step runs (fused steps + integrator) and the one-handle multi-GPU front end
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_driver.py tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
