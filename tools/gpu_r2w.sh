#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/gpu_sphstats.py gas16m > gpurun_out/r2w_sphstats.log 2>&1
tail -3 gpurun_out/r2w_sphstats.log
