#!/bin/bash
# 2-GPU call: bound slice results (test) and the multi-rank e2e with them
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bound_slice or staged_device or late_upload or force_path_equals or call_sequences or bound_results" > gpurun_out/r3c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3c_pytest.log
tail -15 gpurun_out/r3c_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3c_bench2_c3.json 2> gpurun_out/r3c_bench2_c3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --workload plummer1m --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3c_bench2_c1.json 2> gpurun_out/r3c_bench2_c1.err
tail -3 gpurun_out/r3c_bench2_c3.err | cut -c1-400
python - <<'P'
import json
for f in ("r3c_bench2_c3","r3c_bench2_c1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["multi_gpu_check"][:50])
    except Exception as e: print(f, "failed", e)
P
