#!/bin/bash
# 2-GPU call: agb_multi with the NVLink broadcast of the hand-over; timing of the multi handle with pageable / pinned host arrays
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_driver.py -m gpu -q -k "multi or several or two_gpu" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2q_pytest.log
timeout 300 python - > gpurun_out/r2q_multi2.log 2>&1 <<'P'
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
for rep in range(3):
    t0 = time.perf_counter()
    one.set_particles(dict(p)); one.force_path(want["R"] / 1e5, mh, 0.0, e0, 0.5); got = one.results()
    print("one context, disk4m, pageable host arrays in and out: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
one.close()
m = pkg.MultiContext([0, 1], 8)
m.set_particles(dict(p)); m.force_path(want["R"] / 1e5, mh, 0.0, e0, 0.5)
for rep in range(3):
    t0 = time.perf_counter()
    m.set_particles(dict(p)); m.force_path(want["R"] / 1e5, mh, 0.0, e0, 0.5); got = m.results()
    t1 = time.perf_counter()
    print("agb_multi 2 devices, disk4m, pageable host arrays in and out: %.1f ms" % ((t1 - t0) * 1e3), "bitwise equal to one GPU:", all(np.array_equal(got[k], want[k]) for k in got))
m.close()
P
tail -4 gpurun_out/r2q_pytest.log; cat gpurun_out/r2q_multi2.log | tail -8
