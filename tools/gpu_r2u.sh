#!/bin/bash
# GPU call: equal-mass tie sums without sorting, batched child loads in the upward pass, warp-aggregated key histograms
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2u_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2u_bench_c3.json 2> gpurun_out/r2u_bench_c3.err
timeout 300 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2u_bench_c1.json 2> gpurun_out/r2u_bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2u_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2u_ncu1.log 2>&1
tail -4 gpurun_out/r2u_pytest.log
python - <<'P'
import json
for f in ("r2u_bench_c3","r2u_bench_c1"):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()})
P
