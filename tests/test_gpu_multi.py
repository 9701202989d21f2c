"""Two GPUs, two processes (torchrun + NCCL): the target-sharded run reproduces the single-GPU results bit for bit.
bench.py asserts it once per run before timing (`multi_gpu_check`); this test drives that path on a small gas disk.
Skipped on boxes with a single GPU (the driver's 1/2/4/8-GPU scaling run exercises the same assertion at full size)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_slices_equal_single_gpu_bitwise():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29613",
           os.path.join(ROOT, "bench.py"), "--gpus", "2", "--workload", "disk400k", "--steps", "2", "--warmup", "1", "--no-cpu-baseline", "--no-fp64"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["n_gpus"] == 2 and "bit for bit" in (line["multi_gpu_check"] or "")
    assert line["e2e"]["d2h_bytes_per_step"] == 400000 * (4 + 8 * 9)            # every result column of every particle comes back once


def _devices():
    """Two devices when the box has them; otherwise the same device twice (its two contexts then take turns): the slice
    exchange, the row-sharded download and the replicated integrator are the same code either way."""
    import torch
    return [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]


@pytest.mark.parametrize("mixed", [True, False], ids=["mixed", "fp64"])
def test_multi_handle_equals_one_context_bitwise(pkg, mixed):
    """agb_multi_* (one process, several devices, peer-to-peer slice exchange) against a plain one-GPU context: every result
    column bit for bit, through the four calls and through agb_multi_force_path (twice: the second one takes the late-upload path)."""
    import numpy as np
    rng = np.random.default_rng(5)
    p = pkg.ics.disk_galaxy(70000, seed=33)
    n = len(p["x"])
    for k in ("rho", "P", "T", "h", "dUdt", "ax", "ay", "az"):
        p[k] = rng.random(n) + 0.5
    p["next_time"] = np.where(rng.random(n) < 0.7, 0.0, 1e13)              # some particles are not force targets this step
    mh = pkg.ics.gas_mass_in_h(p, 64)
    one = pkg.Context(0, 8)
    one.set_option(pkg.capi.AGB_OPT_PRECISION, 1 if mixed else 0)
    want, _ = pkg.run_step(dict(p), 0.5, 1e18, mh, 0.0, context=one)
    want["visualDensity"] = want["vis"]
    one.close()
    m = pkg.MultiContext(_devices(), 8)
    try:
        m.set_option(pkg.capi.AGB_OPT_PRECISION, 1 if mixed else 0)
        m.set_particles(dict(p))
        R = m.build_tree(); m.visual_density(R / 100000); m.gas_density(mh); m.forces(0.0, 1e18, 0.5)
        assert R == want["R"]
        got = m.results()
        for k in got:
            assert np.array_equal(got[k], want[k]), k
        for rep in range(2):
            m.set_particles(dict(p))
            assert m.force_path(R / 100000, mh, 0.0, 1e18, 0.5) == want["R"]
            got = m.results()
            for k in got:
                assert np.array_equal(got[k], want[k]), (k, rep)
    finally:
        m.close()


def test_multi_handle_resident_loop_equals_one_context(pkg):
    """Device-resident KDK steps with the integrator replicated on every device: same trajectory bits as one context."""
    import numpy as np
    p = pkg.ics.plummer(20000, seed=8, gas_fraction=0.3)
    mh = pkg.ics.gas_mass_in_h(p, 16)

    def run(ctx):
        ctx.set_particles(dict(p))
        ctx.integrator_init(0.02, 1e10, 1e13, 70.0, 1e18)
        R = ctx.build_tree(); ctx.visual_density(R / 100000); ctx.gas_density(mh); ctx.forces(0.0, 1e18, 0.5)
        ctx.integrator_assign_all()
        for _ in range(4):
            t = ctx.step_begin()
            ctx.force_path(R / 100000, mh, t, 1e18, 0.5)
            ctx.step_end()
        st = ctx.state()
        st.update(ctx.results())
        return st
    one = pkg.Context(0, 8)
    want = run(one)
    one.close()
    m = pkg.MultiContext(_devices(), 8)
    try:
        got = run(m)
    finally:
        m.close()
    assert len(np.unique(want["timeStep"])) > 1
    for k in want:
        assert np.array_equal(got[k], want[k]), k
