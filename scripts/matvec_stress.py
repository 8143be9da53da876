"""Stress test of the TMA/DMMA block matvec: repeats W = A X on uploaded matrices and reports, per shape and
variant (DAV_MATVEC_SCHEDULE, DAV_MATVEC_BK), how many runs differ from numpy and WHERE (rows / columns of the wrong
entries, size of the error), to localise intermittent faults.  DAV_B200_LIB selects another build of the library."""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
from fortran_davidson_b200 import davidson as dv

ap = argparse.ArgumentParser()
ap.add_argument("--shapes", default="2048x128,4097x33,3000x64,2500x80,1024x8,2000x20")
ap.add_argument("--variants", default="0:16,1:16,2:16,0:32,2:32")
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
print("library:", fd._lib.LIB_PATH, flush=True)
total_bad = 0
for shp in a.shapes.split(","):
    n, b = [int(x) for x in shp.split("x")]
    rng = np.random.default_rng(n * 1000 + b)
    A = np.asfortranarray(np.diag(np.arange(1.0, n + 1)) + 0.01 * rng.standard_normal((n, n)))
    X = rng.standard_normal((n, b))
    ref = A @ X
    scale = np.abs(ref).max()
    s = fd.DavidsonSolver()
    s.upload(0, A)
    s.set_matvec_impl(dv.MATVEC_TMA_DMMA)
    for var in a.variants.split(","):
        sch, bk = var.split(":")
        os.environ["DAV_MATVEC_SCHEDULE"] = sch
        os.environ["DAV_MATVEC_BK"] = bk
        nbad = 0
        first = None
        for r in range(a.reps):
            W = s.block_matvec(0, X)
            err = np.abs(W - ref)
            bad = np.argwhere(err > 1e-10 * scale)
            if bad.size:
                nbad += 1
                if first is None:
                    rows, cols = np.unique(bad[:, 0]), np.unique(bad[:, 1])
                    first = "rep %d: %d bad entries, rows %d..%d (%d distinct), cols %d..%d (%d distinct), max err %.3e" % (
                        r, len(bad), rows.min(), rows.max(), len(rows), cols.min(), cols.max(), len(cols), err.max())
        total_bad += nbad
        print("n %5d b %3d schedule %s bk %s : %d / %d runs wrong%s" % (n, b, sch, bk, nbad, a.reps,
                                                                       ("   first: " + first) if first else ""), flush=True)
    s.close()
print("TOTAL wrong runs:", total_bad)
