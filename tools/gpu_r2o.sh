#!/bin/bash
# GPU call: walk variants that leave more of the 228 KB to L1 (smaller lists / stacks, fewer warps per CTA)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in H I J K; do
  export AGB200_LIB=$GRAFT_REPO_ROOT/dev_libs/libagb200_$v.so
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2o_bench_c3_$v.json 2> gpurun_out/r2o_bench_c3_$v.err
  timeout 300 python bench.py --workload plummer1m --steps 20 --warmup 3 --no-cpu-baseline --no-fp64 > gpurun_out/r2o_bench_c1_$v.json 2> gpurun_out/r2o_bench_c1_$v.err
done
python - <<'P'
import json
for v in ("H","I","J","K"):
    for w in ("c3","c1"):
        try:
            d=json.loads(open("gpurun_out/r2o_bench_%s_%s.json"%(w,v)).read().strip().splitlines()[-1])
            print(v,w,round(d["ms_per_step"],3),"walk",round(d["roofline"]["kernel_ms"]["k_walk"],3),"spills",d["divergence_counters"]["walk_stack_spills"])
        except Exception as e: print(v,w,"failed",e)
P
