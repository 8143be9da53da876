"""A/B of the block-matvec work schedules (DAV_MATVEC_SCHEDULE = 0 stream-K | 1 waves + stream-K remainder |
2 waves + aligned split-K remainder) and stage depths (DAV_MATVEC_BK = 16 | 32) on ONE resident matrix:
CUDA-event time of the kernel (+ its pack / fixup passes) per launch, median of --reps, as GB/s of the
algorithmic bytes 8 n^2 + 16 n b and TFLOP/s of 2 n^2 b.  Writes a JSON summary when --out is given."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000)
ap.add_argument("--widths", default="16,32,64")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--schedules", default="0,1,2")
ap.add_argument("--bks", default="16,32")
ap.add_argument("--out", default=None)
a = ap.parse_args()
s = fd.DavidsonSolver()
s.generate_diagonal_dominant(0, a.n, 1e-4, None, 0)
rows = []
for b in [int(x) for x in a.widths.split(",")]:
    for bk in [int(x) for x in a.bks.split(",")]:
        for sch in [int(x) for x in a.schedules.split(",")]:
            os.environ["DAV_MATVEC_SCHEDULE"] = str(sch)
            os.environ["DAV_MATVEC_BK"] = str(bk)
            ms = s.bench_block_matvec(0, b, a.reps)
            m = float(np.median(ms))
            r = {"n": a.n, "b": b, "bk": bk, "schedule": sch, "ms": m, "ms_min": float(min(ms)),
                 "GBps": (8.0 * a.n * a.n + 16.0 * a.n * b) / m * 1e-6, "TFLOPs": 2.0 * a.n * a.n * b / m * 1e-9}
            rows.append(r)
            print("n %d b %3d bk %2d schedule %d  ms %.3f (min %.3f)  GB/s %.1f  TF/s %.2f" %
                  (a.n, b, bk, sch, m, r["ms_min"], r["GBps"], r["TFLOPs"]), flush=True)
s.close()
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)
