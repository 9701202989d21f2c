#!/bin/bash
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or against_oracle or C3_gas16m-mixed or C2_disk4m-mixed or dudt or counter_mode or late_upload" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2z_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2z_bench_c3.json 2> gpurun_out/r2z_bench_c3.err
tail -3 gpurun_out/r2z_pytest.log
python - <<'P'
import json
for f in ("r2z_bench_c3",):
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],2), {k: round(v,3) for k,v in d["roofline"]["kernel_ms"].items()}, "records", d.get("sph_records_per_step_rank0"))
P
