#!/bin/bash
# sanitizer pass over the new code paths
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2x_memcheck.log 2>&1; echo "memcheck rc $?" >> gpurun_out/r2x_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2x_racecheck.log 2>&1; echo "racecheck rc $?" >> gpurun_out/r2x_racecheck.log
tail -6 gpurun_out/r2x_memcheck.log; tail -6 gpurun_out/r2x_racecheck.log
