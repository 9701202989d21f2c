#!/bin/bash
# full ncu capture of the final k_walk / k_sph / gather kernels on C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_sph" -c 2 -o gpurun_out/r3r_walk_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r3r_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/r3r_ncu2.log | cut -c1-200
