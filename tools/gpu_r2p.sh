#!/bin/bash
# 8-GPU call: default bench at N = 8 and 4 (as the driver's scaling run launches it), multi-device handle on 8 devices
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
AGB_BENCH_BREAKDOWN=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench8_c3.json 2> gpurun_out/r2p_bench8_c3.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2p_bench4_c3.json 2> gpurun_out/r2p_bench4_c3.err
timeout 300 python - > gpurun_out/r2p_multi8.log 2>&1 <<'P'
import sys, time
sys.path.insert(0, ".")
import numpy as np
import __graft_entry__ as ge
import bench
pkg = ge.load_package()
p, e0, mh, desc = bench.make_particles(pkg, "disk4m")
one = pkg.Context(0, 8)
want, _ = pkg.run_step(dict(p), 0.5, e0, mh, 0.0, context=one)
want["visualDensity"] = want["vis"]
one.close()
m = pkg.MultiContext(list(range(8)), 8)
m.set_particles(dict(p)); m.force_path(want["R"] / 1e5, mh, 0.0, e0, 0.5)
for rep in range(3):
    t0 = time.perf_counter()
    m.set_particles(dict(p)); m.force_path(want["R"] / 1e5, mh, 0.0, e0, 0.5); got = m.results()
    t1 = time.perf_counter()
    print("agb_multi 8 devices, disk4m, host arrays in and out: %.1f ms" % ((t1 - t0) * 1e3), "bitwise equal to one GPU:", all(np.array_equal(got[k], want[k]) for k in got))
m.close()
P
tail -c 2500 gpurun_out/r2p_bench8_c3.err | grep -E "rank 0 breakdown" | cut -c1-700; head -c 700 gpurun_out/r2p_bench8_c3.json; echo; head -c 300 gpurun_out/r2p_bench4_c3.json; echo; cat gpurun_out/r2p_multi8.log | tail -5
