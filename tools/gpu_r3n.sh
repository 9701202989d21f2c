#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AGB_TIMELINE=1 AGB_BENCH_BREAKDOWN=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 --no-fp64 > gpurun_out/r3n_bench2.json 2> gpurun_out/r3n_bench2.err
grep "timeline, bound slice 0\|rank 0 e2e" gpurun_out/r3n_bench2.err | tail -6
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3n_bench2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
P
