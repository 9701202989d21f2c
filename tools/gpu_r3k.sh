#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or late or sequence or staged or bound or devices or gather or slice" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-fp64 > gpurun_out/r3k_bench2.json 2> gpurun_out/r3k_bench2.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3k_bench2.json").read().strip().splitlines()[-1])
print("N=2", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
P
