"""Builds astrogenesis2.0_b200/libagb200.so (the C-ABI library of include/agb200.h) with nvcc for sm_100a.

In-tree, explicit nvcc: the .so travels to the GPU box with the repo snapshot."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libagb200.so")
SOURCES = ["agb_api.cu", "agb_build.cu", "agb_density.cu", "agb_walk.cu", "agb_integrate.cu", "agb_multi.cu", "agb_extended.cu"]
# tree geometry, moments and densities are compared with the reference's separately rounded sums: no FMA contraction there
NO_FMAD = ("agb_build.cu", "agb_density.cu", "agb_integrate.cu")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--use_fast_math=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "agb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + (["-Xptxas", "-v"] if verbose else []) + \
              (["-fmad=false"] if src in NO_FMAD else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-lcudart", "-lpthread", "-ldl"])
    return LIB


if __name__ == "__main__":
    import sys
    print(build_lib(force="-f" in sys.argv, verbose="-v" in sys.argv))
