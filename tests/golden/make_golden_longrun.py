"""Generates tests/golden/longrun_n20k_oracle.json: the CPU oracle on the collapse-exercising long run that
bench.py reports as iterations/s (generate_diagonal_dominant(20000, 5e-2), seed 0, lowest 10, DPR, tol 1e-8,
max_dim_sub = 30: the basis goes 20, 40, 20, 40, ... and needs 20 outer iterations).

    python tests/golden/make_golden_longrun.py            (~35 s on 8 cores, 3.2 GB)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

N, SPARSITY, L, MAX_DIM, TOL = 20000, 5e-2, 10, 30, 1e-8


def main():
    orc.set_num_threads(os.cpu_count() or 1)
    A = orc.generate_diagonal_dominant(N, SPARSITY, None, 0)
    t = time.time()
    r = orc.generalized_eigensolver(A, L, "DPR", 1000, TOL, MAX_DIM)
    dt = time.time() - t
    doc = {"n": N, "sparsity": SPARSITY, "lowest": L, "max_dim_sub": MAX_DIM, "tolerance": TOL, "seed": 0,
           "iters": int(r.iters), "trace_k": [int(x) for x in r.trace_k], "trace_err": [float(x) for x in r.trace_err],
           "eigenvalues": [float(x) for x in r.eigenvalues], "solve_s": dt}
    json.dump(doc, open(os.path.join(ROOT, "tests", "golden", "longrun_n20k_oracle.json"), "w"), indent=1)
    print(json.dumps({k: doc[k] for k in ("n", "iters", "solve_s")}))


if __name__ == "__main__":
    main()
