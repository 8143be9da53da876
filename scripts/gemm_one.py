"""One shape of the tall-skinny GEMM, for ncu:  python scripts/gemm_one.py T 192 64 12500 [partials]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fortran_davidson_b200._lib import check, lib  # noqa: E402

ta, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = 3
ms = (C.c_float * reps)()
check(lib().dav_debug_gemm_bench(C.c_char(ta.encode()), C.c_int64(m), C.c_int64(n), C.c_int64(k), C.c_int(reps),
                                 C.c_int(1 if ta == "T" else 0), ms, None))
print(list(ms))
