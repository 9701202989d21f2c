#!/bin/bash
# GPU call: walk statistics in shared memory (no spills at 2 CTAs/SM); variants with 20 / 24 warps per SM; k_sph capture (2nd launch)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or against_oracle or counter_mode or C3_gas16m or dudt or tiny or inactive or extreme or small_opening" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2n_pytest.log
for v in base F G; do
  if [ $v != base ]; then export AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2n_bench_c3_$v.json 2> gpurun_out/r2n_bench_c3_$v.err
  timeout 300 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2n_bench_c1_$v.json 2> gpurun_out/r2n_bench_c1_$v.err
done
unset AGB200_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sph" -s 1 -c 1 -o gpurun_out/r2n_sph_c3 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fp64 > gpurun_out/r2n_ncu2.log 2>&1
tail -3 gpurun_out/r2n_pytest.log
python - <<'P'
import json
for v in ("base","F","G"):
    for w in ("c3","c1"):
        try:
            d=json.loads(open("gpurun_out/r2n_bench_%s_%s.json"%(w,v)).read().strip().splitlines()[-1])
            print(v,w,round(d["ms_per_step"],3),"walk",round(d["roofline"]["kernel_ms"]["k_walk"],3),"sph",round(d["roofline"]["kernel_ms"]["k_sph"],3))
        except Exception as e: print(v,w,"failed",e)
P
