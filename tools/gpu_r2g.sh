#!/bin/bash
# 2-GPU call: the multi-device handle over real peer-to-peer copies, the torchrun 2-rank bitwise check
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_box.txt; nvidia-smi topo -m >> gpurun_out/r2g_box.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py -m gpu -q -k "multi or several or two_gpu" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2g_pytest.log
AGB_BENCH_BREAKDOWN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 5 --warmup 2 --no-cpu-baseline --no-fp64 > gpurun_out/r2g_bench2_c3.json 2> gpurun_out/r2g_bench2_c3.err
tail -8 gpurun_out/r2g_pytest.log; tail -c 1500 gpurun_out/r2g_bench2_c3.err; head -c 600 gpurun_out/r2g_bench2_c3.json
