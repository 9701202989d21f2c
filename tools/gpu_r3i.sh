#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AGB_TIMELINE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
grep "agb timeline" gpurun_out/r3i_bench.err | tail -4
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3i_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2))
P
