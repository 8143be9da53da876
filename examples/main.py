#!/usr/bin/env python
"""The reference's demo program src/main.f90:31-74 against the B200 library (same calls and printed checks):
generalized dense problem, dim = 100, lowest 3, GJD and DPR, max_dim_sub = 10, tolerance 1e-5."""
import _common  # noqa: F401
from fortran_davidson_b200 import generalized_eigensolver, generate_diagonal_dominant
from fortran_davidson_b200.array_utils import norm

dim = 100
mtx = generate_diagonal_dominant(dim, 1e-3)
stx = generate_diagonal_dominant(dim, 1e-3, 1.0, seed=1)

eigenvalues_GJD, eigenvectors_GJD, iter_i = generalized_eigensolver(mtx, 3, "GJD", 100, 1e-5, 10, stx)
print(" GJD algorithm converged in: ", iter_i, " iterations!")
eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(mtx, 3, "DPR", 100, 1e-5, 10, stx)
print(" DPR algorithm converged in: ", iter_i, " iterations!")

print(" Test 1")
test_norm_eigenvalues = norm(eigenvalues_GJD - eigenvalues_DPR)
print(" Check that eigenvalues norm computed by different methods are the same: ", test_norm_eigenvalues < 1e-6)

print(" Test 2")
print(" Check that eigenvalue equation:  H V = l S V  holds!")
for name, ev, vec in (("DPR", eigenvalues_DPR, eigenvectors_DPR), ("GJD", eigenvalues_GJD, eigenvectors_GJD)):
    print(" %s method:" % name)
    for j in range(3):
        xs = mtx @ vec[:, j] - ev[j] * (stx @ vec[:, j])
        print(" eigenvalue ", j + 1, ": ", ev[j], "||Error||: ", norm(xs))
