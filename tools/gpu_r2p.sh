#!/bin/bash
# 8-GPU call: default bench at N = 8 and 4 (as the driver's scaling run launches it), multi-device handle on 8 devices
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AGB_BENCH_BREAKDOWN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench8_c3.json 2> gpurun_out/r2p_bench8_c3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2p_bench4_c3.json 2> gpurun_out/r2p_bench4_c3.err

tail -c 2500 gpurun_out/r2p_bench8_c3.err | grep -E "rank 0 breakdown" | cut -c1-700; head -c 700 gpurun_out/r2p_bench8_c3.json; echo; head -c 300 gpurun_out/r2p_bench4_c3.json; echo; 
