#!/bin/bash
# GPU call: full suite with the extended mode in the library; extended bench lines on C1 / C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extended.py -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2s_pytest.log

timeout 900 python bench.py --extended --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2s_bench_c3_ext.json 2> gpurun_out/r2s_bench_c3_ext.err

tail -8 gpurun_out/r2s_pytest.log
python - <<'P'
import json
for f in ("r2s_bench_c1_ext","r2s_bench_c3_ext","r2s_bench_c3"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items() if k in ("build","gas_density","k_far","k_walk","k_sph")}, "frac", round(d["roofline"]["frac"],4), "inter/target", d["interactions_per_s"])
    except Exception as e: print(f, "failed", e); print(open("gpurun_out/%s.err"%f).read()[-800:])
P
