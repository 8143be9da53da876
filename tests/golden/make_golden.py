"""Generates the committed golden fixtures.  Run HERE (needs /root/reference for matrix.txt):

    python tests/golden/make_golden.py

* matrix_100.npy      : the reference's only deterministic input, src/tests/matrix.txt (100x100,
                        row-major text), converted losslessly to float64 .npy.
* golden_cases.json   : for every named case the oracle's eigenvalues / iteration count / basis
                        schedule / residual trace, plus scipy.linalg.eigh's lowest eigenvalues of
                        the same matrix (the reference's own acceptance check,
                        src/tests/test_davidson.py:36-40,67-69).
The GPU parity tests replay the same seeded inputs and compare against these numbers.
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sl

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as orc  # noqa: E402

REF_MATRIX = "/root/reference/src/tests/matrix.txt"


def main():
    m = np.loadtxt(REF_MATRIX)
    assert m.shape == (100, 100)
    np.save(os.path.join(HERE, "matrix_100.npy"), m)

    cases = {}

    def dense_case(name, A, lowest, method, max_it, tol, max_dim, B=None, desc=None):
        r = orc.generalized_eigensolver(A, lowest, method, max_it, tol, max_dim, B)
        es = sl.eigh(A, B)[0][:lowest]
        cases[name] = dict(desc=desc, lowest=lowest, method=method, max_iterations=max_it, tolerance=tol,
                           max_dim_sub=max_dim, iters=int(r.iters), trace_k=[int(x) for x in r.trace_k],
                           trace_err=[float(x) for x in r.trace_err], eigenvalues=[float(x) for x in r.eigenvalues],
                           eigh=[float(x) for x in es])

    for method in ("DPR", "GJD"):
        dense_case("matrix_txt_" + method, m, 3, method, 1000, 1e-8, None,
                   desc="src/tests/matrix.txt, L=3, default max_dim")
        A = orc.generate_diagonal_dominant(50, 1e-4, seed=0)
        B = orc.generate_diagonal_dominant(50, 1e-4, 1.0, seed=1)
        dense_case("readme_std_" + method, A, 3, method, 1000, 1e-8, 20,
                   desc="README config: generate_diagonal_dominant(50,1e-4) seed 0")
        dense_case("readme_gev_" + method, A, 3, method, 1000, 1e-8, 20, B,
                   desc="README config + second_matrix generate_diagonal_dominant(50,1e-4,1.0) seed 1")
        A = orc.generate_diagonal_dominant(50, 1e-3, seed=2)
        B = orc.generate_diagonal_dominant(50, 1e-3, 1.0, seed=3)
        dense_case("test_dense_numpy_std_" + method, A, 3, method, 1000, 1e-8, None,
                   desc="test_dense_numpy.f90:17-21")
        dense_case("test_dense_numpy_gen_" + method, A, 3, method, 1000, 1e-8, 10, B,
                   desc="test_dense_numpy.f90:30-32")
        A = orc.generate_diagonal_dominant(100, 1e-3, seed=4)
        B = orc.generate_diagonal_dominant(100, 1e-3, 1.0, seed=5)
        dense_case("main_f90_" + method, A, 3, method, 100, 1e-5, 10, B, desc="main.f90:49-54")
    A = orc.generate_diagonal_dominant(1000, 1e-2, seed=0)
    dense_case("collapse_n1000_DPR", A, 3, "DPR", 1000, 1e-10, 10, desc="exercises the collapse branch (6,12,6,...)")
    B = orc.generate_diagonal_dominant(1000, 1e-2, 1.0, seed=1)
    dense_case("collapse_n1000_gev_DPR", A, 3, "DPR", 1000, 1e-10, 10, B, desc="collapse branch, generalized")
    A = orc.generate_diagonal_dominant(2000, 5e-2, seed=0)
    dense_case("collapse_n2000_DPR", A, 10, "DPR", 1000, 1e-8, 100, desc="20,40,80,160,20,40")
    A = orc.generate_diagonal_dominant(400, 1e-3, seed=7)
    dense_case("notconverged_DPR", A, 4, "DPR", 2, 1e-14, None, desc="max_iterations hit: iters = max_iterations+1")

    def free_case(name, dim, op_a, op_b, lowest, max_it, tol, max_dim, desc):
        r = orc.generalized_eigensolver_free(dim, op_a, op_b, lowest, "DPR", max_it, tol, max_dim)
        Ma, Mb = orc.operator_matrix(op_a, dim), orc.operator_matrix(op_b, dim)
        es = sl.eigh(Ma, Mb)[0][:lowest]
        cases[name] = dict(desc=desc, dim=dim, op_a=op_a, op_b=op_b, lowest=lowest, method="DPR",
                           max_iterations=max_it, tolerance=tol, max_dim_sub=max_dim, iters=int(r.iters),
                           trace_k=[int(x) for x in r.trace_k], trace_err=[float(x) for x in r.trace_err],
                           eigenvalues=[float(x) for x in r.eigenvalues], eigh=[float(x) for x in es])

    free_case("free_test_50", 50, orc.OP_TEST_MTX, orc.OP_TEST_STX, 3, 1000, 1e-8, 20,
              "test_free_numpy.f90 / test_free_properties.f90: on-the-fly 50x50 mtx/stx")
    free_case("free_benchmark_1000", 1000, orc.OP_BENCHMARK_MTX, orc.OP_IDENTITY, 3, 1000, 1e-8, 20,
              "benchmark_free.f90:88-102: dim 1000, stx = identity")
    free_case("free_benchmark_300_L8", 300, orc.OP_BENCHMARK_MTX, orc.OP_IDENTITY, 8, 1000, 1e-8, None,
              "benchmark operator, default max_dim")

    json.dump(cases, open(os.path.join(HERE, "golden_cases.json"), "w"), indent=1)
    for k, v in cases.items():
        print(k, v["iters"], v["trace_k"], v["eigenvalues"][:3])


if __name__ == "__main__":
    main()
