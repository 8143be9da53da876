"""Matrix-free benchmark operator at n = 200,000 on one GPU: iterations, solve time, basis schedule and how often the
fast block orthonormalisation was rejected (the configs[4] regression check of r02: nearly parallel DPR corrections)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fortran_davidson_b200 as fd
s = fd.DavidsonSolver()
n = 200000
s.set_operator(0, n, fd.OP_BENCHMARK_MTX); s.set_operator(1, n, fd.OP_IDENTITY)
for _ in range(2):
    ev, vec, it = s.solve(32, "DPR", 1000, 1e-8, None, want_vectors=True)
    st = s.stats()
    print("free n=%d iters=%s solve_ms=%.1f matvec_launches=%d pip_fallbacks=%d trace=%s" % (n, it, st.solve_ms, st.matvec_launches, st.pip_fallbacks, list(st.trace_k[:st.trace_len])))
