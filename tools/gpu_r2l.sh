#!/bin/bash
# GPU call: fast 21-level keys, flag-gated upward reads — tests, full default bench (as the driver runs it), C1, reference arm, launch list + full captures on C3
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=4 > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2l_pytest.log
timeout 900 python bench.py > gpurun_out/r2l_bench_c3.json 2> gpurun_out/r2l_bench_c3.err
timeout 300 python bench.py --workload plummer1m --no-cpu-baseline > gpurun_out/r2l_bench_c1.json 2> gpurun_out/r2l_bench_c1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2l_launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2l_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_walk|k_sph" -c 2 -o gpurun_out/r2l_walk_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2l_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_keygen|k_sort_onesweep|k_gather|k_links|k_upward_levels|k_gas_fold|k_visual" -c 16 -o gpurun_out/r2l_build_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2l_ncu3.log 2>&1
tail -4 gpurun_out/r2l_pytest.log; head -c 400 gpurun_out/r2l_bench_c3.json; echo; cat gpurun_out/r2l_bench_ref.json | head -c 1500
