#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(nvidia-smi topo -m; lscpu | grep -i "numa\|socket\|^CPU(s)"; nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null) > gpurun_out/r3l_topo.txt 2>&1
for aff in 0 1; do
AGB_BENCH_NO_AFFINITY=$aff timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-fp64 > gpurun_out/r3l_bench2_$aff.json 2> gpurun_out/r3l_bench2_$aff.err
done
python - <<'P'
import json
for a in (0, 1):
    d=json.loads(open("gpurun_out/r3l_bench2_%d.json"%a).read().strip().splitlines()[-1])
    print("no_affinity=%d"%a, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"].get("host_numa_node_rank0"))
P
cat gpurun_out/r3l_topo.txt | head -30
