#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slice_densities or fp32_range or sequence or blown or late_upload" 2>&1 | tail -15
