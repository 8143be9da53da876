"""Event timings of the tall-skinny GEMM kernel (csrc/dgemm.cu) on the shapes of the headline solve, for one rank of
1 and of 8 GPUs, across the warp-group settings (DAV_GEMM_KG) and split targets.  usage: python scripts/gemm_bench.py"""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortran_davidson_b200._lib import check, lib  # noqa: E402


def bench(ta, m, n, k, reps=20, partials=0):
    ms = (C.c_float * reps)()
    err = C.c_double(0.0)
    check(lib().dav_debug_gemm_bench(C.c_char(ta.encode()), C.c_int64(m), C.c_int64(n), C.c_int64(k), C.c_int(reps),
                                     C.c_int(partials), ms, C.byref(err)))
    return float(np.median(list(ms))) * 1e3, err.value


def main():
    shapes = []
    for nl in (12500, 100000):
        shapes += [("T", 64, 32, nl), ("T", 128, 64, nl), ("T", 192, 64, nl), ("T", 256, 128, nl)]
    if "--nn" in sys.argv:
        for nl in (12500, 100000):
            shapes += [("N", nl, 16, 32), ("N", nl, 16, 128), ("N", nl, 48, 64), ("N", nl, 112, 128),
                       ("N", nl, 32, 64), ("N", nl, 64, 128), ("N", nl, 64, 192)]
    variants = [("impl0", {"DAV_GEMM_IMPL": "0"}), ("kg1", {"DAV_GEMM_KG": "1"}), ("kg2", {"DAV_GEMM_KG": "2"}),
                ("kg4", {"DAV_GEMM_KG": "4"}), ("auto", {}), ("novec", {"DAV_GEMM_VEC": "0"}),
                ("novec_kg2", {"DAV_GEMM_VEC": "0", "DAV_GEMM_KG": "2"}),
                ("t148", {"DAV_GEMM_SPLIT_TARGET": "148"}), ("t444", {"DAV_GEMM_SPLIT_TARGET": "444"})]
    print("%-28s" % "shape (ta, m, n, k)" + "".join("%11s" % v[0] for v in variants) + "   [us, median of 20]")
    for (ta, m, n, k) in shapes:
        row = []
        for name, env in variants:
            for kk in ("DAV_GEMM_IMPL", "DAV_GEMM_KG", "DAV_GEMM_SPLIT_TARGET", "DAV_GEMM_VEC"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            us, err = bench(ta, m, n, k, partials=1 if ta == "T" else 0)
            row.append(us)
        flops = 2.0 * m * n * k
        print("%-28s" % str((ta, m, n, k)) + "".join("%11.1f" % r for r in row) + "   %.1f GF" % (flops * 1e-9))


if __name__ == "__main__":
    main()
