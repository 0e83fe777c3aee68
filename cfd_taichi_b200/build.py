"""Builds libsph_b200.so (the C-ABI of include/sph_b200.h) in-tree with nvcc for sm_100a.

The sweep translation unit is compiled twice: strict fp32 (-fmad=false, bit-exact against the
oracle) and fast (FMA contraction + approximate reciprocals, same neighbour sets).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("SPH_BUILD_DIR") or os.path.join(HERE, "_build")  # SPH_BUILD_DIR: variant builds for A/B runs
LIB = os.path.join(OUT, "libsph_b200.so")

def _nccl_include():
    # the torch-bundled NCCL (the library the process maps at run time); fall back to the system header
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            p = os.path.join(list(spec.submodule_search_locations)[0], "include")
            if os.path.exists(os.path.join(p, "nccl.h")):
                return p
    except Exception:
        pass
    return "/usr/include"


NCCL_INC = _nccl_include()

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math"] + ARCH


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


DEBUG_OUT = os.path.join(HERE, "_build_debug")
DEBUG_LIB = os.path.join(DEBUG_OUT, "libsph_b200.so")


def build_debug(force=False):
    """The same library with -DSPH_DEBUG_BOUNDS=1 (every neighbour-list entry, list length and candidate segment is
    validated before use, SPH_ERR_BOUNDS latched): what tests/test_gpu_bounds.py runs instead of compute-sanitizer,
    which is closed on this GPU pool.  Loaded through SPH_B200_LIB; never the default."""
    global OUT, LIB
    saved = (OUT, LIB, os.environ.get("SPH_EXTRA_NVCC"))
    OUT, LIB = DEBUG_OUT, DEBUG_LIB
    os.environ["SPH_EXTRA_NVCC"] = ((saved[2] or "") + " -DSPH_DEBUG_BOUNDS=1").strip()
    try:
        return build(force=force)
    finally:
        OUT, LIB = saved[0], saved[1]
        if saved[2] is None:
            os.environ.pop("SPH_EXTRA_NVCC", None)
        else:
            os.environ["SPH_EXTRA_NVCC"] = saved[2]


def build(force=False, verbose=False):
    os.makedirs(OUT, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(HERE, "..", "include", "sph_b200.h"))
    extra = os.environ.get("SPH_EXTRA_NVCC", "").split()
    fast_extra = os.environ.get("SPH_FAST_NVCC", "").split()     # A/B knob: extra flags of the fast sweep unit only
    units = [
        ("sph_grid.cu", "sph_grid.o", extra),
        ("sph_api.cu", "sph_api.o", extra),
        ("sph_render.cu", "sph_render.o", extra),
        ("sph_multigpu.cu", "sph_multigpu.o", ["-I", NCCL_INC] + extra),
        ("sph_sweeps.cu", "sph_sweeps_strict.o", ["-DSPH_STRICT=1", "-fmad=false"] + extra),
        # fast kernels: FMA contraction, approximate division / square root (2 ulp: inside the 1e-5 budget, asserted
        # sweep by sweep in tests/test_gpu_fast_parity.py); the cull and the cell hash use explicit IEEE intrinsics
        ("sph_sweeps.cu", "sph_sweeps_fast.o", ["-DSPH_STRICT=0", "-fmad=true"] + (
            [] if os.environ.get("SPH_FAST_IEEE_DIV") else ["-prec-div=false", "-prec-sqrt=false"]) + fast_extra + extra),
    ]
    objs = []
    procs = []
    for src, obj, extra in units:
        srcp, objp = os.path.join(CSRC, src), os.path.join(OUT, obj)
        objs.append(objp)
        if force or _newer([srcp] + hdrs, objp):
            cmd = ["nvcc"] + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", objp]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(" ".join(cmd))
            print(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or _newer(objs, LIB):
        cmd = ["nvcc", "-shared", "-o", LIB] + objs + ARCH + ["-cudart", "shared", "-ldl"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
