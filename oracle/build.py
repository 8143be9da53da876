"""Build recipe for the CPU oracle (test infrastructure, see davidson_oracle.cpp header).

g++ compiles oracle/davidson_oracle.cpp into oracle/_build/liboracle.so and links it against the
real LAPACK/BLAS bundled with scipy (OpenBLAS, LP64, `scipy_`-prefixed symbols).  The reference
itself (Fortran) cannot be compiled in this image: there is no Fortran compiler, so there is no
oracle/_ref/.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "liboracle.so")
SRC = os.path.join(HERE, "davidson_oracle.cpp")


def find_openblas():
    import scipy

    libs = os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs")
    cands = sorted(glob.glob(os.path.join(libs, "libscipy_openblas*.so")))
    if not cands:
        raise RuntimeError("scipy's bundled OpenBLAS not found under %s" % libs)
    return libs, os.path.basename(cands[0])


def build(force=False):
    if (not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC)
            and os.path.getmtime(OUT) >= os.path.getmtime(__file__)):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    libdir, libname = find_openblas()
    cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC,
           "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
