#!/bin/bash
# final-like verification: what the driver runs at round end
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r3z_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3z_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r3z_smoke.log
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r3z_bench_ref.json 2> gpurun_out/r3z_bench_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err
tail -3 gpurun_out/r3z_pytest.log; tail -3 gpurun_out/r3z_smoke.log; tail -4 gpurun_out/r3z_bench_ref.err; tail -4 gpurun_out/r3z_bench.err
