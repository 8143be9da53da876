"""Generates the matrix and runs the block matvec kernel alone for the given widths (target of the
`ncu --set full -k regex:matvec_kernel` capture; also prints event timings when run bare)."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000); ap.add_argument("--widths", default="16,32,64")
ap.add_argument("--reps", type=int, default=1); ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--free", action="store_true", help="benchmark_free on-the-fly operator instead of a dense matrix")
a = ap.parse_args()
s = fd.DavidsonSolver()
s.set_matvec_impl(a.impl)
if a.free:
    s.set_operator(0, a.n, fd.OP_BENCHMARK_MTX)
else:
    s.generate_diagonal_dominant(0, a.n, 1e-4, None, 0)
for b in [int(x) for x in a.widths.split(",")]:
    ms = s.bench_block_matvec(0, b, a.reps)
    m = float(np.median(ms))
    if a.free:
        print("free n %d b %d ms %.3f  Gentries/s %.1f  TF/s(2*n*n*b) %.2f" % (a.n, b, m, 1e-6 * a.n * a.n / m,
                                                                             2.0 * a.n * a.n * b / m * 1e-9), flush=True)
    else:
        print("n %d b %d ms %.3f GB/s %.1f TF/s %.2f" % (a.n, b, m, (8.0 * a.n * a.n + 16.0 * a.n * b) / m * 1e-6,
                                                      2.0 * a.n * a.n * b / m * 1e-9), flush=True)
s.close()
