#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "late or sequence or staged or bound or inactive or dudt or golden" 2>&1 | tail -3
AGB_TIMELINE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err
grep "agb timeline" gpurun_out/r3j_bench.err | tail -2
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 --workload plummer1m > gpurun_out/r3j_bench_c1.json 2> gpurun_out/r3j_bench_c1.err
python - <<'P'
import json
for f in ("r3j_bench", "r3j_bench_c1"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
P
