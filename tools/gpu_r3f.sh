#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bound_slice or call_sequences or late_upload or staged_device" > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3f_pytest.log
tail -12 gpurun_out/r3f_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err
timeout 300 python bench.py --workload plummer1m --steps 20 --no-cpu-baseline --no-fp64 > gpurun_out/r3f_bench_c1.json 2> gpurun_out/r3f_bench_c1.err
tail -3 gpurun_out/r3f_bench.err
python - <<'P'
import json
for f in ("r3f_bench","r3f_bench_c1"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), d["e2e"]["d2h_bytes_per_step"])
    except Exception as e: print(f, "failed", e)
P
