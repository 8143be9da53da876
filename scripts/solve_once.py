"""Runs warm-up + ONE timed solve of a configuration (for ncu launch lists)."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100000); ap.add_argument("--lowest", type=int, default=16)
ap.add_argument("--max-dim", type=int, default=0); ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--method", default="DPR"); ap.add_argument("--gev", action="store_true")
ap.add_argument("--free", action="store_true", help="matrix-free benchmark_free operator + identity")
a = ap.parse_args()
s = fd.DavidsonSolver()
if a.free:
    s.set_operator(0, a.n, fd.OP_BENCHMARK_MTX)
    s.set_operator(1, a.n, fd.OP_IDENTITY)
else:
    s.generate_diagonal_dominant(0, a.n, 1e-4, None, 0)
if a.gev and not a.free:
    s.generate_diagonal_dominant(1, a.n, 1e-4, 1.0, 1)
for _ in range(a.warm + 1):
    ev, vec, it = s.solve(a.lowest, a.method, 1000, 1e-8, a.max_dim or None)
st = s.stats()
print("n", a.n, "lowest", a.lowest, a.method, "gev" if a.gev else "", "free" if a.free else "", "schedule", list(st.trace_k[:st.trace_len]), "ev0", ev[0])
print("iters", it, "solve_ms", st.solve_ms, "matvec_ms", st.matvec_ms, "rr", st.rr_ms, "orth", st.orth_ms, "resid", st.resid_ms, "proj", st.proj_ms, "launches", st.kernel_launches, "gjd_inner", st.gjd_inner_iterations)
