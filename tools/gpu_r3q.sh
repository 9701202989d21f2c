#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or sequence or late or dudt or node_table" 2>&1 | tail -2
for sp in 0 1; do
AGB_GATHER_SPLIT=$sp timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3q_bench_$sp.json 2> gpurun_out/r3q_bench_$sp.err
AGB_GATHER_SPLIT=$sp timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 --workload plummer1m > gpurun_out/r3q_bench_c1_$sp.json 2> gpurun_out/r3q_bench_c1_$sp.err
done
python - <<'P'
import json
for f in ("r3q_bench_0", "r3q_bench_1", "r3q_bench_c1_0", "r3q_bench_c1_1"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items() if k.startswith("build")})
P
