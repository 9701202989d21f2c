#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/debug_hooks.py > gpurun_out/r2j_debug.log 2>&1
grep -v Warning gpurun_out/r2j_debug.log | grep -v "rate =" | tail -20
