#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-fp64 > gpurun_out/r3v_bench2.json 2> gpurun_out/r3v_bench2.err
python - <<'P'
import json
for f in ("r3v_bench2",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d.get("multi_gpu_check"))
    except Exception as e:
        print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-1500:])
P
