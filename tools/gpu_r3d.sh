#!/bin/bash
# final checks on 1 GPU: full suite, default bench (masses follow the positions in the host hand-over)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3d_pytest.log
timeout 900 python bench.py > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
timeout 300 python bench.py --workload plummer1m --no-cpu-baseline > gpurun_out/r3d_bench_c1.json 2> gpurun_out/r3d_bench_c1.err
tail -3 gpurun_out/r3d_pytest.log
python - <<'P'
import json
for f in ("r3d_bench","r3d_bench_c1"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), "fp64", round(d["fp64"]["ms_per_step"],2), "resident", d["resident_sim_step"])
P
