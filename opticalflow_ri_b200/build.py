"""Build libofri.so (sm_100a) in-tree with nvcc.  `python -m opticalflow_ri_b200.build` or build().

The library is CUDA-only: there is no CPU build of the product.  nvcc cross-compiles on a box without a GPU."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libofri.so")
SOURCES = ["ofri_api.cu", "ofri_stages.cu", "ofri_hs.cu", "ofri_hs_tma.cu", "ofri_ls.cu", "ofri_ls_tma.cu", "ofri_comm.cu", "ofri_farneback.cu", "ofri_lk.cu"]
HEADERS = ["ofri_internal.h", "ofri_pixel.cuh", "ofri_hs_common.cuh", "ofri_ls_common.cuh", "ofri_tma.cuh", "ofri_tables.h", "ofri_spline.cuh", os.path.join("..", "..", "include", "ofri.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "--cudart", "static"]


def _nvcc():
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libofri.so cannot be built (there is no CPU fallback)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, phase_timing=False):
    """phase_timing: the profiling variant libofri_phase.so (-DOFRI_PHASE_TIMING: per-phase clock64() accumulators in the
    persistent sweep kernels; tools/phase_timing.py) -- never loaded by the product path."""
    lib = os.path.join(HERE, "libofri_phase.so") if phase_timing else LIB
    if not force and not phase_timing and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", "phase") if phase_timing else os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-DOFRI_PHASE_TIMING"] + os.environ.get("OFRI_EXTRA_DEFS", "").split() if phase_timing else []) + \
            (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [nvcc, "-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, phase_timing="--phase" in sys.argv))
