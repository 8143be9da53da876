"""One-GPU check of the block matvec on the rectangular row blocks of the sharded solve (m = rows of a rank, k = n):
TMA/DMMA kernel under each schedule against the library's tall-skinny GEMM, data generated on the device."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
from fortran_davidson_b200._lib import check
L = fd.lib()
L.dav_debug_matvec_rect.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
shapes = [(50000, 100000, 64), (25000, 100000, 64), (12544, 100000, 64), (50000, 100000, 32), (25000, 100000, 128),
          (12192, 100000, 16), (19000, 30000, 100)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
bad = 0
for m, k, b in shapes:
    for sch in ("default", "0", "1", "2"):
        if sch == "default":
            os.environ.pop("DAV_MATVEC_SCHEDULE", None)
        else:
            os.environ["DAV_MATVEC_SCHEDULE"] = sch
        d, s = C.c_double(), C.c_double()
        check(L.dav_debug_matvec_rect(0, m, k, b, C.byref(d), C.byref(s)))
        ok = d.value <= 1e-11 * s.value
        bad += not ok
        print("m %6d k %6d b %3d schedule %-7s max|diff| %.3e  max|W| %.3e  %s" % (m, k, b, sch, d.value, s.value,
                                                                                  "ok" if ok else "MISMATCH"), flush=True)
print("RECT_CHECK_%s" % ("PASSED" if bad == 0 else "FAILED"))
