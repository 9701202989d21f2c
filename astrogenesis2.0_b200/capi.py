"""ctypes binding of include/agb200.h (libagb200.so).  No fallback: if the library is missing or no
B200 is present, every call raises."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_pd = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8)

AGB_MEM_HOST, AGB_MEM_DEVICE = 0, 1
AGB_OPT_TARGET_COUNTERS = 1
AGB_OPT_PRECISION = 2
AGB_OPT_COOLING = 3
AGB_OPT_STAR_FORMATION = 4
AGB_OPT_EXTENDED = 5
AGB_OPT_SLICE_PIECE = 6
AGB_OPT_SLICE_DENSITIES = 7

EXPORTS = [
    "agb_create", "agb_destroy", "agb_set_particles", "agb_set_particles_staged", "agb_set_particles_aos", "agb_build_tree", "agb_visual_density",
    "agb_gas_density", "agb_forces", "agb_forces_slice", "agb_force_path", "agb_get_slice_count", "agb_get_slice_results", "agb_get_slice_results_all", "agb_bind_slice_results", "agb_get_kernel_ms", "agb_get_results", "agb_bind_results", "agb_get_results_aos", "agb_get_counters",
    "agb_set_option", "agb_get_tree_particles", "agb_get_node_count", "agb_get_nodes", "agb_get_target_counters",
    "agb_get_phase_ms", "agb_get_stream", "agb_get_launch_count", "agb_microbench",
    "agb_integrator_init", "agb_integrator_assign_all", "agb_step_begin", "agb_step_end", "agb_get_state", "agb_get_subgrid_state", "agb_strerror", "agb_last_error", "agb_version",
    "agb_multi_create", "agb_multi_destroy", "agb_multi_device_count", "agb_multi_context", "agb_multi_set_option", "agb_multi_set_particles",
    "agb_multi_set_particles_aos", "agb_multi_build_tree", "agb_multi_visual_density", "agb_multi_gas_density", "agb_multi_forces", "agb_multi_force_path",
    "agb_multi_get_results", "agb_multi_get_results_aos", "agb_multi_integrator_init", "agb_multi_integrator_assign_all", "agb_multi_step_begin",
    "agb_multi_step_end", "agb_multi_get_state", "agb_multi_get_subgrid_state", "agb_multi_last_error",
]


class Particles(C.Structure):
    _fields_ = [("n", C.c_int64)] + [(k, _pd) for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "U", "next_time", "mu")] + \
        [("type", _pu8)] + [(k, _pd) for k in ("rho", "P", "T", "h", "dUdt", "ax", "ay", "az")]


class Results(C.Structure):
    _fields_ = [(k, _pd) for k in ("ax", "ay", "az", "dUdt", "h", "rho", "P", "T", "visualDensity")]


class AosLayout(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("position", "velocity", "acc", "mass", "type", "U", "next_time", "mu", "rho", "P", "T", "h", "dUdt", "visualDensity")]


COUNTER_FIELDS = ("n_particles", "n_in_tree", "n_outliers", "n_nodes", "n_active", "max_depth", "edge_dropped", "interactions",
                  "node_interactions", "leaf_interactions", "sph_interactions", "node_visits", "mac_exact_fallbacks",
                  "groups", "gas_groups", "gas_orphans", "gas_ties_exact", "gas_ties_unresolved",
                  "walk_rounds", "walk_popped", "walk_straddling", "walk_opened", "walk_tiles", "walk_stack_spills",
                  "walk_ent_wide", "walk_ent_half", "walk_ent_quarter", "walk_bits_wide", "walk_bits_half", "walk_bits_quarter", "walk_ent_far",
                  "walk_ent_class0", "walk_ent_class1", "walk_ent_class2", "sph_records")


class Counters(C.Structure):
    _fields_ = [(k, C.c_int64) for k in COUNTER_FIELDS]


class AgbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("agb200 status %d: %s" % (status, msg))
        self.status = status


_lib = None


def library_path():
    return _build.LIB


def load(build_if_needed=True):
    """Load libagb200.so (building it in-tree with nvcc when stale). Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("AGB200_LIB") or _build.LIB          # A/B runs of two builds of the same ABI (development only)
    if path == _build.LIB and build_if_needed and _build.needs_build():
        _build.build_lib()
    if not os.path.exists(path):
        raise RuntimeError("libagb200.so is not built; run __graft_entry__.build() (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.agb_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
    lib.agb_destroy.argtypes = [vp]
    lib.agb_set_particles.argtypes = [vp, C.POINTER(Particles), C.c_int]
    lib.agb_set_particles_aos.argtypes = [vp, C.POINTER(vp), C.c_int64, C.POINTER(AosLayout)]
    if hasattr(lib, "agb_set_particles_staged"):
        lib.agb_set_particles_staged.argtypes = [vp, C.POINTER(Particles), C.c_int, vp, vp, vp]
    lib.agb_build_tree.argtypes = [vp, _pd]
    lib.agb_visual_density.argtypes = [vp, C.c_double]
    lib.agb_gas_density.argtypes = [vp, C.c_double]
    lib.agb_forces.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.agb_forces_slice.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int]
    if hasattr(lib, "agb_force_path"):                        # absent only in older builds loaded through AGB200_LIB
        lib.agb_force_path.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, _pd]
    lib.agb_get_slice_count.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    lib.agb_get_slice_results.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint32)] + [_pd] * 4 + [C.c_int]
    lib.agb_get_slice_results_all.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(Results), C.c_int]
    if hasattr(lib, "agb_bind_slice_results"):
        lib.agb_bind_slice_results.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(Results)]
    lib.agb_get_kernel_ms.argtypes = [vp, _pd]
    lib.agb_get_results.argtypes = [vp, C.POINTER(Results), C.c_int]
    lib.agb_bind_results.argtypes = [vp, C.POINTER(Results), C.c_int]
    lib.agb_get_results_aos.argtypes = [vp, C.POINTER(vp), C.c_int64, C.POINTER(AosLayout)]
    lib.agb_get_counters.argtypes = [vp, C.POINTER(Counters)]
    lib.agb_set_option.argtypes = [vp, C.c_int, C.c_int64]
    lib.agb_get_tree_particles.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.agb_get_node_count.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.agb_get_nodes.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)] + [_pd] * 8
    lib.agb_get_target_counters.argtypes = [vp] + [C.POINTER(C.c_int32)] * 4
    lib.agb_get_phase_ms.argtypes = [vp, _pd]
    lib.agb_get_stream.argtypes = [vp, C.POINTER(vp)]
    lib.agb_get_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.agb_microbench.argtypes = [vp, C.c_int, _pd]
    lib.agb_integrator_init.argtypes = [vp] + [C.c_double] * 5
    lib.agb_integrator_assign_all.argtypes = [vp]
    lib.agb_step_begin.argtypes = [vp, _pd]
    lib.agb_step_end.argtypes = [vp]
    lib.agb_get_state.argtypes = [vp] + [_pd] * 9
    if hasattr(lib, "agb_get_subgrid_state"):
        lib.agb_get_subgrid_state.argtypes = [vp, _pu8, _pd]
    lib.agb_strerror.restype = C.c_char_p
    lib.agb_strerror.argtypes = [C.c_int]
    lib.agb_last_error.restype = C.c_char_p
    lib.agb_last_error.argtypes = [vp]
    lib.agb_version.restype = C.c_char_p
    if hasattr(lib, "agb_multi_create"):
        lib.agb_multi_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int, C.c_int]
        lib.agb_multi_destroy.argtypes = [vp]
        lib.agb_multi_device_count.argtypes = [vp]
        lib.agb_multi_context.argtypes = [vp, C.c_int, C.POINTER(vp)]
        lib.agb_multi_set_option.argtypes = [vp, C.c_int, C.c_int64]
        lib.agb_multi_set_particles.argtypes = [vp, C.POINTER(Particles)]
        lib.agb_multi_set_particles_aos.argtypes = [vp, C.POINTER(vp), C.c_int64, C.POINTER(AosLayout)]
        lib.agb_multi_build_tree.argtypes = [vp, _pd]
        lib.agb_multi_visual_density.argtypes = [vp, C.c_double]
        lib.agb_multi_gas_density.argtypes = [vp, C.c_double]
        lib.agb_multi_forces.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        lib.agb_multi_force_path.argtypes = [vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _pd]
        lib.agb_multi_get_results.argtypes = [vp, C.POINTER(Results)]
        lib.agb_multi_get_results_aos.argtypes = [vp, C.POINTER(vp), C.c_int64, C.POINTER(AosLayout)]
        lib.agb_multi_integrator_init.argtypes = [vp] + [C.c_double] * 5
        lib.agb_multi_integrator_assign_all.argtypes = [vp]
        lib.agb_multi_step_begin.argtypes = [vp, _pd]
        lib.agb_multi_step_end.argtypes = [vp]
        lib.agb_multi_get_state.argtypes = [vp] + [_pd] * 9
        lib.agb_multi_get_subgrid_state.argtypes = [vp, _pu8, _pd]
        lib.agb_multi_last_error.restype = C.c_char_p
        lib.agb_multi_last_error.argtypes = [vp]
    _lib = lib
    return lib


def check(ctx, status):
    if status != 0:
        lib = load()
        msg = lib.agb_strerror(status).decode()
        if ctx:
            extra = lib.agb_last_error(ctx).decode()
            if extra:
                msg += " (" + extra + ")"
        raise AgbError(status, msg)


def dptr(a, ctype=C.c_double):
    """Pointer to a numpy array's data, or NULL."""
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))
