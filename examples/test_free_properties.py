#!/usr/bin/env python
"""The reference's src/tests/test_free_properties.f90 against the B200 library: matrix-free solve of the 50 x 50
on-the-fly test operators through host callbacks, residuals H V = l B V printed like the reference does."""
import sys

import _common  # noqa: F401
import numpy as np

from fortran_davidson_b200 import generalized_eigensolver
from fortran_davidson_b200.array_utils import diagonal, norm
from fortran_davidson_b200.test_utils import (apply_mtx_to_vect, apply_stx_to_vect, compute_matrix_on_the_fly,
                                              compute_stx_on_the_fly)

dim = 50
mtx = np.zeros((dim, dim), order="F")
stx = np.zeros((dim, dim), order="F")
for j in range(1, dim + 1):
    mtx[:, j - 1] = compute_matrix_on_the_fly(j, dim)
    stx[:, j - 1] = compute_stx_on_the_fly(j, dim)

eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(
    apply_mtx_to_vect, 3, "DPR", 1000, 1e-8, 20, fun_second_matrix_gemv=apply_stx_to_vect, dim=dim)

print("eigenvalues: " + "".join("%8.4f" % e for e in eigenvalues_DPR))
print(" Test 1")
print(" Check that eigenvalue equation:  H V = l B V holds")
print(" DPR method:")
ok = True
for j in range(3):
    xs = mtx @ eigenvectors_DPR[:, j] - eigenvalues_DPR[j] * (stx @ eigenvectors_DPR[:, j])
    flag = norm(xs) < 1e-8
    ok &= bool(flag)
    print("error: %10.3e" % norm(xs))
    print("eigenvalue %2d: %12.5e succeeded: %s" % (j + 1, eigenvalues_DPR[j], "T" if flag else "F"))

print(" Test 2")
print(" If V are the eigenvector then V * V^T = I")
zs = diagonal(eigenvectors_DPR @ eigenvectors_DPR.T)
print("DPR method: %s" % ("T" if norm(zs[:3]) < np.sqrt(3.0) else "F"))
sys.exit(0 if ok else 1)
