#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r3t_bench4.json 2> gpurun_out/r3t_bench4.err
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r3t_bench4.json").read().strip().splitlines()[-1])
    print("N=4", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d.get("multi_gpu_check"), d.get("fp64", {}) and d["fp64"].get("ms_per_step"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/r3t_bench4.err").read()[-2000:])
P
