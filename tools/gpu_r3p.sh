#!/bin/bash
# launch list of the final round-2 state (default bench command, 2 steps) + timeline line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3p_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r3p_ncu1.log 2>&1
tail -2 gpurun_out/r3p_ncu1.log | cut -c1-300
wc -l gpurun_out/r3p_launches_c3.csv
