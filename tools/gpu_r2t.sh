#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_walk_ext|k_ext_density" -c 2 -o gpurun_out/r2t_ext -f python bench.py --extended --workload disk400k --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log
