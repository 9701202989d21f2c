#!/bin/bash
# verification of the final state: full GPU suite, smoke, default bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3s_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3s_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r3s_smoke.log
timeout 900 python bench.py > gpurun_out/r3s_bench.json 2> gpurun_out/r3s_bench.err
tail -3 gpurun_out/r3s_pytest.log; tail -1 gpurun_out/r3s_smoke.log
python - <<'P'
import json
d=json.loads(open("gpurun_out/r3s_bench.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
P
