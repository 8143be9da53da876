#!/usr/bin/env python
"""The reference's src/tests/test_call_lapack.f90 against the device mirrors of `lapack_wrapper`: dumps
test_lapack_*.txt for the checks of src/tests/test_lapack.py:14-66."""
import _common  # noqa: F401
from fortran_davidson_b200 import generate_diagonal_dominant
from fortran_davidson_b200.lapack_wrapper import lapack_generalized_eigensolver, lapack_qr
from fortran_davidson_b200.test_utils import write_matrix, write_vector

dim = 50
mtx = generate_diagonal_dominant(dim, 1e-3)
copy = mtx.copy(order="F")
stx = generate_diagonal_dominant(dim, 1e-3, seed=1)
write_matrix("test_lapack_matrix.txt", mtx)
write_matrix("test_lapack_stx.txt", stx)

# standard eigenvalue problem
eigenvalues, eigenvectors = lapack_generalized_eigensolver(copy)
write_vector("test_lapack_eigenvalues.txt", eigenvalues)
write_matrix("test_lapack_eigenvectors.txt", eigenvectors)

# General eigenvalue problem
eigenvalues, eigenvectors = lapack_generalized_eigensolver(copy, stx)
write_vector("test_lapack_eigenvalues_gen.txt", eigenvalues)
write_matrix("test_lapack_eigenvectors_gen.txt", eigenvectors)

# Lapack orthonormalization
write_matrix("test_lapack_qr.txt", lapack_qr(mtx))
