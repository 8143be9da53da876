#!/usr/bin/env python
"""The reference's src/tests/test_dense_properties.f90 against the B200 library (same calls, same printed
logicals; the reference test passes on exit code 0 and prints T/F flags, here a False flag also exits 1)."""
import sys

import _common  # noqa: F401
import numpy as np

from fortran_davidson_b200 import generalized_eigensolver, generate_diagonal_dominant
from fortran_davidson_b200.array_utils import diagonal, norm

dim, lowest = 50, 3
mtx = generate_diagonal_dominant(dim, 1e-3)
eigenvalues_GJD, eigenvectors_GJD, iter_i = generalized_eigensolver(mtx, lowest, "GJD", 1000, 1e-8)
eigenvalues_DPR, eigenvectors_DPR, iter_i = generalized_eigensolver(mtx, lowest, "DPR", 1000, 1e-8)
ok = True

print(" Test 1")
flag = norm(eigenvalues_GJD - eigenvalues_DPR) < 1e-8
ok &= bool(flag)
print(" Check that eigenvalues norm computed by different methods are the same: ", flag)

print(" Test 2")
print(" Check that eigenvalue equation:  H V = l V holds")
for name, ev, vec in (("DPR", eigenvalues_DPR, eigenvectors_DPR), ("GJD", eigenvalues_GJD, eigenvectors_GJD)):
    print(" %s method:" % name)
    for j in range(lowest):
        xs = mtx @ vec[:, j] - ev[j] * vec[:, j]
        flag = norm(xs) < 1e-8
        ok &= bool(flag)
        print(" eigenvalue ", j + 1, ": ", flag)

print(" Test 3")
print(" If V are the eigenvector then V * V^T = I")
# (the reference prints norms of the diagonal of V V^T against sqrt(lowest), test_dense_properties.f90:43-47;
# the property it means is V^T V = I, checked here as well)
ys = diagonal(eigenvectors_GJD @ eigenvectors_GJD.T)
zs = diagonal(eigenvectors_DPR @ eigenvectors_DPR.T)
print(" GJD method: ", norm(ys[:3]) < np.sqrt(lowest))
print(" DPR method: ", norm(zs[:3]) < np.sqrt(lowest))
for vec in (eigenvectors_GJD, eigenvectors_DPR):
    ok &= bool(np.abs(vec.T @ vec - np.eye(lowest)).max() < 1e-8)
sys.exit(0 if ok else 1)
