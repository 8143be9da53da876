#!/usr/bin/env python
"""The reference's src/tests/test_dense_numpy.f90 against the B200 library: same calls, same dump files
(test_dense_spec_*.txt, test_dense_gen_*.txt) for the checks of src/tests/test_davidson.py:15-51.
Writes into the current directory; prints nothing on stderr when all is well (the driver treats any stderr
output as a failure, test_davidson.py:90-92)."""
import _common  # noqa: F401
from fortran_davidson_b200 import generalized_eigensolver, generate_diagonal_dominant
from fortran_davidson_b200.test_utils import write_matrix, write_vector

dim, lowest = 50, 3
mtx = generate_diagonal_dominant(dim, 1e-3)

# call eigenvalue solver (test_dense_numpy.f90:21-22)
eigenvalues_GJD, eigenvectors_GJD, iter_i = generalized_eigensolver(mtx, lowest, "GJD", 1000, 1e-8)
eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(mtx, lowest, "DPR", 1000, 1e-8)

write_matrix("test_dense_spec_matrix.txt", mtx)
write_vector("test_dense_spec_eigenvalues_GJD.txt", eigenvalues_GJD)
write_vector("test_dense_spec_eigenvalues_DPR.txt", eigenvalues_DPR)
write_matrix("test_dense_spec_eigenvectors_GJD.txt", eigenvectors_GJD)
write_matrix("test_dense_spec_eigenvectors_DPR.txt", eigenvectors_DPR)

# call generalized eigenvalue solver (test_dense_numpy.f90:31-33)
stx = generate_diagonal_dominant(dim, 1e-3, 1.0, seed=1)
eigenvalues_GJD_gen, eigenvectors_GJD_gen, iter_i = generalized_eigensolver(mtx, lowest, "GJD", 1000, 1e-8, 10, stx)
eigenvalues_DPR_gen, eigenvectors_DPR_gen, iter_i = generalized_eigensolver(mtx, lowest, "DPR", 1000, 1e-8, 10, stx)

write_matrix("test_dense_gen_matrix.txt", mtx)
write_matrix("test_dense_gen_stx.txt", stx)
write_vector("test_dense_gen_eigenvalues_GJD.txt", eigenvalues_GJD_gen)
write_vector("test_dense_gen_eigenvalues_DPR.txt", eigenvalues_DPR_gen)
write_matrix("test_dense_gen_eigenvectors_GJD.txt", eigenvectors_GJD_gen)
write_matrix("test_dense_gen_eigenvectors_DPR.txt", eigenvectors_DPR_gen)
