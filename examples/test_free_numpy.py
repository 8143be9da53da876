#!/usr/bin/env python
"""The reference's src/tests/test_free_numpy.f90 against the B200 library: the matrix-free solver driven by the
two operator callbacks apply_mtx_to_vect / apply_stx_to_vect, dumps matrix_free.txt, stx_free.txt,
eigenvalues_DPR_free.txt, eigenvectors_DPR_free.txt for src/tests/test_davidson.py:54-79."""
import _common  # noqa: F401
import numpy as np

from fortran_davidson_b200 import generalized_eigensolver
from fortran_davidson_b200.test_utils import (apply_mtx_to_vect, apply_stx_to_vect, compute_matrix_on_the_fly,
                                              compute_stx_on_the_fly, write_matrix, write_vector)

dim, lowest = 50, 3
mtx = np.zeros((dim, dim), order="F")
stx = np.zeros((dim, dim), order="F")
for j in range(1, dim + 1):
    mtx[:, j - 1] = compute_matrix_on_the_fly(j, dim)
    stx[:, j - 1] = compute_stx_on_the_fly(j, dim)

# Write matrices down to test the eigenvalues against numpy (test_free_numpy.f90:24-26)
write_matrix("matrix_free.txt", mtx)
write_matrix("stx_free.txt", stx)

eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(
    apply_mtx_to_vect, lowest, "DPR", 1000, 1e-8, 20, fun_second_matrix_gemv=apply_stx_to_vect, dim=dim)

write_vector("eigenvalues_DPR_free.txt", eigenvalues_DPR)
write_matrix("eigenvectors_DPR_free.txt", eigenvectors_DPR)
