"""Builds libdavidson_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m fortran_davidson_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  No torch, no cuBLAS:
the only runtime dependencies are the CUDA runtime (linked statically) and, for multi-GPU,
libnccl.so.2 resolved with dlopen at run time.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libdavidson_b200.so")
SOURCES = ["capi.cu", "solver.cu", "comm.cu", "dgemm.cu", "smalldense.cu", "trideig.cu", "vecops.cu", "freeops.cu", "freeops_dmma.cu",
           "matvec_dmma.cu", "gjd.cu", "microbench.cu", "densesolve.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = (["-DDAV_TRIDIAG_PROFILE=" + os.environ["DAV_TRIDIAG_PROFILE"]] if os.environ.get("DAV_TRIDIAG_PROFILE") else []) + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "--extended-lambda"]


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "davidson_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(path)
            and os.path.getmtime(obj) >= _deps_mtime()):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), SOURCES))
    objs = [o for o, _ in results]
    logs = "".join(l for _, l in results)
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart",
                                                     "static", "-ldl", "-lpthread", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print(logs)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
