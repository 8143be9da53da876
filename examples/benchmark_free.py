#!/usr/bin/env python
"""The reference's src/benchmark_free.f90:80-111 against the B200 library: matrix-free solve of the on-the-fly
operator (cos(log(sqrt(atan2))) entries, overlap = identity), lowest 3, DPR, max_dim_sub 20, tolerance 1e-8.
The operators are the built-in device generators DAV_OP_BENCHMARK_MTX / DAV_OP_IDENTITY (entries generated on the
fly in the kernel, nothing dim x dim is stored on the device); --callbacks drives the same solve through host
callbacks like the Fortran program does (mtx_gemv / stx_gemv = free_matmul of the generator).
    python examples/benchmark_free.py [dim] [--callbacks]"""
import sys

import _common  # noqa: F401
import numpy as np

from fortran_davidson_b200 import OP_BENCHMARK_MTX, OP_IDENTITY, free_matmul, generalized_eigensolver
from fortran_davidson_b200.array_utils import norm
from fortran_davidson_b200.davidson import compute_matrix_on_the_fly, generalized_eigensolver_builtin

args = [a for a in sys.argv[1:] if not a.startswith("--")]
dim = int(args[0]) if args else 1000
lowest = 3

mtx = np.zeros((dim, dim), order="F")
stx = np.zeros((dim, dim), order="F")
for j in range(1, dim + 1):
    mtx[:, j - 1] = compute_matrix_on_the_fly(OP_BENCHMARK_MTX, j, dim)
    stx[:, j - 1] = compute_matrix_on_the_fly(OP_IDENTITY, j, dim)

if "--callbacks" in sys.argv:
    eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(
        lambda x: free_matmul(OP_BENCHMARK_MTX, x), lowest, "DPR", 1000, 1e-8, 20,
        fun_second_matrix_gemv=lambda x: free_matmul(OP_IDENTITY, x), dim=dim)
else:
    eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver_builtin(
        dim, OP_BENCHMARK_MTX, OP_IDENTITY, lowest, "DPR", 1000, 1e-8, 20)

for j in range(lowest):
    xs = mtx @ eigenvectors_DPR[:, j] - eigenvalues_DPR[j] * (stx @ eigenvectors_DPR[:, j])
    print(" error: ", norm(xs))
    print(" eigenvalue ", j + 1, ": ", eigenvalues_DPR[j], " succeeded: ", norm(xs) < 1e-8)
