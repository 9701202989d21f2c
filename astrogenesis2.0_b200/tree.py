"""Host-side mirror of the reference's force-path interface.

The reference reaches the path through `class Tree` (simulation/src/Physics/Tree/Tree.h:13-27):
`Tree(Simulation*)`, `buildTree()`, `calcVisualDensity()`, `calcGasDensity()`, `calculateForces()`
and the public `root->radius`; the parameters are public fields of `Simulation`
(Simulation.h:33-70: numberOfParticles, theta, e0, massInH, globalTime, visualDensityRadius).
`Simulation` and `Tree` below keep those names and meanings; particles are a structure-of-arrays
dict (keys as in agb_particles) instead of `std::vector<Particle*>`.  Everything runs on the GPU
through the C ABI; there is no CPU path.
"""
import ctypes as C

import numpy as np

from . import capi

_F8 = ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu", "rho", "P", "T", "h", "dUdt", "ax", "ay", "az")
_OUT = ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T", "visualDensity")


class Simulation:
    """The fields of the reference's Simulation that the path reads (Simulation.h:33-70)."""

    def __init__(self, particles, theta=0.5, e0=1e19, massInH=1e40, globalTime=0.0):
        self.particles = particles
        self.numberOfParticles = len(particles["x"])
        self.theta = theta
        self.e0 = e0
        self.massInH = massInH
        self.globalTime = globalTime
        self.visualDensityRadius = 0.0


class _Root:
    radius = 0.0


class Context:
    """Persistent device context (pooled memory); one per GPU, reused by successive Trees."""

    def __init__(self, device=0, compat_cores=8):
        self.lib = capi.load()
        self.h = C.c_void_p()
        capi.check(None, self.lib.agb_create(C.byref(self.h), int(device), int(compat_cores)))
        self._keep = None

    def close(self):
        if self.h:
            self.lib.agb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- hand-over
    def set_particles(self, p):
        n = len(p["x"])
        st = capi.Particles()
        st.n = n
        keep = {}
        for k in _F8:
            a = p.get(k)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                keep[k] = a
            setattr(st, k, capi.dptr(a))
        t = np.ascontiguousarray(p["type"], dtype=np.uint8)
        keep["type"] = t
        st.type = capi.dptr(t, C.c_uint8)
        capi.check(self.h, self.lib.agb_set_particles(self.h, C.byref(st), capi.AGB_MEM_HOST))
        self._keep = keep          # uploads are asynchronous: the arrays must outlive agb_build_tree
        self.n = n

    def set_particles_device(self, ptrs, n, events=None):
        """ptrs: dict name -> device address (int) of float64 arrays (uint8 for 'type'); read in place.
        events: optional (ready_positions, ready_next_time, ready_all) cudaEvent_t handles (ints, 0 = ready now) recorded behind
        whatever still produces the arrays on the caller's streams (agb_set_particles_staged)."""
        st = capi.Particles()
        st.n = n
        for k in _F8:
            setattr(st, k, C.cast(C.c_void_p(ptrs.get(k, 0) or None), C.POINTER(C.c_double)))
        st.type = C.cast(C.c_void_p(ptrs["type"]), C.POINTER(C.c_uint8))
        if events is None:
            capi.check(self.h, self.lib.agb_set_particles(self.h, C.byref(st), capi.AGB_MEM_DEVICE))
        else:
            capi.check(self.h, self.lib.agb_set_particles_staged(self.h, C.byref(st), capi.AGB_MEM_DEVICE, *[C.c_void_p(int(e) or None) for e in events]))
        self.n = n

    # ---- the four calls
    def build_tree(self):
        r = C.c_double()
        capi.check(self.h, self.lib.agb_build_tree(self.h, C.byref(r)))
        return r.value

    def visual_density(self, radius):
        capi.check(self.h, self.lib.agb_visual_density(self.h, float(radius)))

    def gas_density(self, massInH):
        capi.check(self.h, self.lib.agb_gas_density(self.h, float(massInH)))

    def forces(self, globalTime, e0, theta, part=0, nparts=1):
        capi.check(self.h, self.lib.agb_forces_slice(self.h, float(globalTime), float(e0), float(theta), int(part), int(nparts)))

    def force_path(self, visual_radius, massInH, globalTime, e0, theta, part=0, nparts=1):
        """build_tree + visual_density + gas_density + forces with one host synchronisation; returns this step's root radius."""
        r = C.c_double()
        capi.check(self.h, self.lib.agb_force_path(self.h, float(visual_radius), float(massInH), float(globalTime), float(e0), float(theta),
                                                   int(part), int(nparts), C.byref(r)))
        return r.value

    # ---- results
    def results(self, names=_OUT):
        out = {k: np.empty(self.n) for k in names}
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, capi.dptr(out.get(k)))
        capi.check(self.h, self.lib.agb_get_results(self.h, C.byref(r), capi.AGB_MEM_HOST))
        return out

    def slice_count(self, part, nparts):
        c = C.c_int64()
        capi.check(self.h, self.lib.agb_get_slice_count(self.h, int(part), int(nparts), C.byref(c)))
        return c.value

    def slice_results(self, part, nparts, names=("ax", "ay", "az", "dUdt"), out=None):
        """Compact results of one target slice (agb_get_slice_results): dict with `index` (caller-order positions of the
        slice's targets, tree order) and the requested columns.  `out` may hold preallocated (pinned) numpy arrays."""
        cnt = self.slice_count(part, nparts)
        res = {"index": (out["index"][:cnt] if out else np.empty(cnt, np.uint32))}
        for k in names:
            res[k] = out[k][:cnt] if out else np.empty(cnt)
        r = capi.Results()
        for k in _OUT:                                     # any of the nine result columns (agb_get_slice_results_all)
            setattr(r, k, capi.dptr(res.get(k)))
        capi.check(self.h, self.lib.agb_get_slice_results_all(self.h, int(part), int(nparts), capi.dptr(res["index"], C.c_uint32), C.byref(r), capi.AGB_MEM_HOST))
        return res

    def bind_slice_results(self, part, nparts, out):
        """Register (pinned) host arrays — `index` (uint32) and any of the nine result columns, each with room for the slice — as the
        destination of slice (part, nparts): force_path(.., part, nparts) then delivers them itself (agb_bind_slice_results).
        out = None unbinds."""
        if out is None:
            capi.check(self.h, self.lib.agb_bind_slice_results(self.h, 0, 1, None, None))
            return
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, capi.dptr(out.get(k)))
        capi.check(self.h, self.lib.agb_bind_slice_results(self.h, int(part), int(nparts), capi.dptr(out["index"], C.c_uint32), C.byref(r)))
        self._bound_slice = out

    def bind_results(self, out):
        """Register host arrays (dict name -> float64 numpy array of length n, ideally pinned) as the destination of the
        results: density outputs are sent while the walk runs (agb_bind_results).  `results_into(out)` completes them."""
        if out is None:
            capi.check(self.h, self.lib.agb_bind_results(self.h, None, capi.AGB_MEM_HOST))
            return
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, capi.dptr(out.get(k)))
        capi.check(self.h, self.lib.agb_bind_results(self.h, C.byref(r), capi.AGB_MEM_HOST))

    def results_into(self, out):
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, capi.dptr(out.get(k)))
        capi.check(self.h, self.lib.agb_get_results(self.h, C.byref(r), capi.AGB_MEM_HOST))
        return out

    def results_device(self, ptrs):
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, C.cast(C.c_void_p(ptrs.get(k, 0) or None), C.POINTER(C.c_double)))
        capi.check(self.h, self.lib.agb_get_results(self.h, C.byref(r), capi.AGB_MEM_DEVICE))

    def counters(self):
        c = capi.Counters()
        capi.check(self.h, self.lib.agb_get_counters(self.h, C.byref(c)))
        return {k: getattr(c, k) for k in capi.COUNTER_FIELDS}

    def set_option(self, opt, value):
        capi.check(self.h, self.lib.agb_set_option(self.h, int(opt), int(value)))

    def tree_particles(self):
        ld = np.empty(self.n, np.int32); hi = np.empty(self.n, np.uint64); lo = np.empty(self.n, np.uint64)
        capi.check(self.h, self.lib.agb_get_tree_particles(self.h, capi.dptr(ld, C.c_int32), capi.dptr(hi, C.c_uint64), capi.dptr(lo, C.c_uint64)))
        return ld, hi, lo

    def nodes(self):
        m = C.c_int64()
        capi.check(self.h, self.lib.agb_get_node_count(self.h, C.byref(m)))
        m = m.value
        nd = {"depth": np.empty(m, np.int32), "count": np.empty(m, np.int64), "dup": np.empty(m, np.int32),
              "key_hi": np.empty(m, np.uint64), "key_lo": np.empty(m, np.uint64)}
        for k in ("mass", "comx", "comy", "comz", "gasMass", "mvx", "mvy", "mvz"):
            nd[k] = np.empty(m)
        capi.check(self.h, self.lib.agb_get_nodes(self.h, capi.dptr(nd["depth"], C.c_int32), capi.dptr(nd["count"], C.c_int64), capi.dptr(nd["dup"], C.c_int32),
                                                  capi.dptr(nd["key_hi"], C.c_uint64), capi.dptr(nd["key_lo"], C.c_uint64),
                                                  *[capi.dptr(nd[k]) for k in ("mass", "comx", "comy", "comz", "gasMass", "mvx", "mvy", "mvz")]))
        return nd

    def target_counters(self):
        a = [np.empty(self.n, np.int32) for _ in range(4)]
        capi.check(self.h, self.lib.agb_get_target_counters(self.h, *[capi.dptr(x, C.c_int32) for x in a]))
        return dict(zip(("visits", "acc_nodes", "acc_leaves", "sph"), a))

    def phase_ms(self):
        a = (C.c_double * 5)()
        capi.check(self.h, self.lib.agb_get_phase_ms(self.h, a))
        return dict(zip(("build", "visual", "gas_density", "walk_kernel", "forces"), list(a)))

    def kernel_ms(self):
        a = (C.c_double * 8)()
        capi.check(self.h, self.lib.agb_get_kernel_ms(self.h, a))
        return dict(zip(("k_far", "k_walk", "k_sph", "build_keys", "build_sort", "build_gather", "build_links", "build_upward"), list(a)))

    def stream(self):
        s = C.c_void_p()
        capi.check(self.h, self.lib.agb_get_stream(self.h, C.byref(s)))
        return s.value or 0

    # ---- device-resident driver loop
    def integrator_init(self, eta, min_ts, max_ts, H0, e0):
        capi.check(self.h, self.lib.agb_integrator_init(self.h, float(eta), float(min_ts), float(max_ts), float(H0), float(e0)))

    def integrator_assign_all(self):
        capi.check(self.h, self.lib.agb_integrator_assign_all(self.h))

    def step_begin(self):
        t = C.c_double()
        capi.check(self.h, self.lib.agb_step_begin(self.h, C.byref(t)))
        return t.value

    def step_end(self):
        capi.check(self.h, self.lib.agb_step_end(self.h))

    def state(self):
        names = ("x", "y", "z", "vx", "vy", "vz", "U", "next_time", "timeStep")
        out = {k: np.empty(self.n) for k in names}
        capi.check(self.h, self.lib.agb_get_state(self.h, *[capi.dptr(out[k]) for k in names]))
        return out

    def subgrid_state(self):
        """(type, sfr) after device-resident steps with AGB_OPT_COOLING / AGB_OPT_STAR_FORMATION."""
        t = np.empty(self.n, np.uint8); s = np.empty(self.n)
        capi.check(self.h, self.lib.agb_get_subgrid_state(self.h, capi.dptr(t, C.c_uint8), capi.dptr(s)))
        return t, s

    def microbench(self, kind):
        """0: FP64 FMA TFLOP/s, 1: FP32 FMA TFLOP/s, 2: HBM copy GB/s (measured, not part of the path)."""
        v = C.c_double()
        capi.check(self.h, self.lib.agb_microbench(self.h, int(kind), C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64()
        capi.check(self.h, self.lib.agb_get_launch_count(self.h, C.byref(v)))
        return v.value


class MultiContext:
    """Several GPUs of one box behind one handle (agb_multi_*): every device builds the same tree, walks a slice of the targets,
    the slices are exchanged device to device.  Results are bit-identical to a one-GPU Context."""

    def __init__(self, devices, compat_cores=8):
        self.lib = capi.load()
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        capi.check(None, self.lib.agb_multi_create(C.byref(self.h), devs, len(devices), int(compat_cores)))
        self._keep = None
        self.n = 0

    def _ck(self, st):
        if st != 0:
            raise capi.AgbError(st, self.lib.agb_strerror(st).decode() + " (" + self.lib.agb_multi_last_error(self.h).decode() + ")")

    def close(self):
        if self.h:
            self.lib.agb_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, opt, value):
        self._ck(self.lib.agb_multi_set_option(self.h, int(opt), int(value)))

    def set_particles(self, p):
        n = len(p["x"])
        st = capi.Particles()
        st.n = n
        keep = {}
        for k in _F8:
            a = p.get(k)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float64)
                keep[k] = a
            setattr(st, k, capi.dptr(a))
        t = np.ascontiguousarray(p["type"], dtype=np.uint8)
        keep["type"] = t
        st.type = capi.dptr(t, C.c_uint8)
        self._ck(self.lib.agb_multi_set_particles(self.h, C.byref(st)))
        self._keep = keep
        self.n = n

    def build_tree(self):
        r = C.c_double()
        self._ck(self.lib.agb_multi_build_tree(self.h, C.byref(r)))
        return r.value

    def visual_density(self, radius):
        self._ck(self.lib.agb_multi_visual_density(self.h, float(radius)))

    def gas_density(self, massInH):
        self._ck(self.lib.agb_multi_gas_density(self.h, float(massInH)))

    def forces(self, globalTime, e0, theta):
        self._ck(self.lib.agb_multi_forces(self.h, float(globalTime), float(e0), float(theta)))

    def force_path(self, visual_radius, massInH, globalTime, e0, theta):
        r = C.c_double()
        self._ck(self.lib.agb_multi_force_path(self.h, float(visual_radius), float(massInH), float(globalTime), float(e0), float(theta), C.byref(r)))
        return r.value

    def results(self, names=_OUT):
        out = {k: np.empty(self.n) for k in names}
        r = capi.Results()
        for k in _OUT:
            setattr(r, k, capi.dptr(out.get(k)))
        self._ck(self.lib.agb_multi_get_results(self.h, C.byref(r)))
        return out

    def integrator_init(self, eta, min_ts, max_ts, H0, e0):
        self._ck(self.lib.agb_multi_integrator_init(self.h, float(eta), float(min_ts), float(max_ts), float(H0), float(e0)))

    def integrator_assign_all(self):
        self._ck(self.lib.agb_multi_integrator_assign_all(self.h))

    def step_begin(self):
        t = C.c_double()
        self._ck(self.lib.agb_multi_step_begin(self.h, C.byref(t)))
        return t.value

    def step_end(self):
        self._ck(self.lib.agb_multi_step_end(self.h))

    def state(self):
        names = ("x", "y", "z", "vx", "vy", "vz", "U", "next_time", "timeStep")
        out = {k: np.empty(self.n) for k in names}
        self._ck(self.lib.agb_multi_get_state(self.h, *[capi.dptr(out[k]) for k in names]))
        return out


class Tree:
    """Drop-in for the reference's Tree (Tree.h:13-27): same construction, same four calls, `root.radius`."""

    def __init__(self, simulation, context=None, device=0, compat_cores=8):
        self.simulation = simulation
        self.ctx = context or Context(device, compat_cores)
        self._own = context is None
        self.root = _Root()

    def buildTree(self):
        self.ctx.set_particles(self.simulation.particles)
        self.root.radius = self.ctx.build_tree()

    def calcVisualDensity(self):
        self.ctx.visual_density(self.simulation.visualDensityRadius)

    def calcGasDensity(self):
        self.ctx.gas_density(self.simulation.massInH)

    def calculateForces(self):
        s = self.simulation
        self.ctx.forces(s.globalTime, s.e0, s.theta)
        # the reference writes into Particle; here the arrays of the SoA dict are replaced
        out = self.ctx.results()
        p = s.particles
        for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T"):
            p[k] = out[k]
        p["visualDensity"] = out["visualDensity"]

    def close(self):
        if self._own:
            self.ctx.close()


def run_step(p, theta, e0, massInH, globalTime=0.0, cores=8, context=None, counters=False):
    """build -> visual density (radius = R/1e5, Simulation.cpp:126) -> gas density -> forces, like
    Simulation::init (Simulation.cpp:120-139). Returns (results dict, Context)."""
    ctx = context or Context(0, cores)
    if counters:
        ctx.set_option(capi.AGB_OPT_TARGET_COUNTERS, 1)
    sim = Simulation(p, theta, e0, massInH, globalTime)
    t = Tree(sim, ctx)
    t.buildTree()
    sim.visualDensityRadius = t.root.radius / 100000
    t.calcVisualDensity()
    t.calcGasDensity()
    t.calculateForces()
    out = {k: sim.particles[k] for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T")}
    out["vis"] = sim.particles["visualDensity"]
    out["R"] = t.root.radius
    return out, ctx
