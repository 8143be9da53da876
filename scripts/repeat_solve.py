"""Repeats a solve several times on a fresh solver and prints iteration count / schedule each time
(used under compute-sanitizer and to check run-to-run reproducibility)."""
import argparse, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2000); ap.add_argument("--lowest", type=int, default=10)
ap.add_argument("--max-dim", type=int, default=100); ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--sparsity", type=float, default=5e-2); ap.add_argument("--method", default="DPR")
ap.add_argument("--gev", action="store_true"); ap.add_argument("--fresh", action="store_true")
a = ap.parse_args()
s = None
for rep in range(a.reps):
    if s is None or a.fresh:
        s = fd.DavidsonSolver()
        s.generate_diagonal_dominant(0, a.n, a.sparsity, None, 0)
        if a.gev:
            s.generate_diagonal_dominant(1, a.n, a.sparsity, 1.0, 1)
    ev, vec, it = s.solve(a.lowest, a.method, 1000, 1e-8, a.max_dim or None)
    st = s.stats()
    print("rep", rep, "iters", it, list(st.trace_k[:st.trace_len]), "err", ["%.2e" % e for e in st.trace_err[:st.trace_len]], "ev0 %.15f" % ev[0], flush=True)
