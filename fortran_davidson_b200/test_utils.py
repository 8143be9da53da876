"""Mirror of the reference's test helper module (src/tests/test_utils.f90): the text dump format its Python
drivers read back (test_davidson.py, test_lapack.py) and the on-the-fly test operators, so the programs under
examples/ produce exactly the files those drivers expect.  No numerics here: the operators are the device
generators DAV_OP_TEST_MTX / DAV_OP_TEST_STX behind `free_matmul` (davidson.f90:526-569)."""
import numpy as np

from . import davidson as _dv


def write_vector(path_file, vector):
    """write_vector (test_utils.f90:138-150): one list-directed value per line."""
    with open(path_file, "w") as fh:
        for v in np.asarray(vector, dtype=np.float64).ravel():
            fh.write("  %.17g\n" % v)


def write_matrix(path_file, mtx):
    """write_matrix (test_utils.f90:153-167): one value per line, ROW-major traversal (i outer, j inner)."""
    m = np.asarray(mtx, dtype=np.float64)
    with open(path_file, "w") as fh:
        for i in range(m.shape[0]):
            for j in range(m.shape[1]):
                fh.write("  %.17g\n" % m[i, j])


def read_matrix(path_file, dim):
    """read_matrix (test_utils.f90:118-135): `dim` rows of `dim` list-directed values (a row may wrap over lines)."""
    vals = np.array(open(path_file).read().split(), dtype=np.float64)
    if vals.size != dim * dim:
        raise ValueError("%s holds %d values, expected %d" % (path_file, vals.size, dim * dim))
    return np.asfortranarray(vals.reshape(dim, dim))


def compute_matrix_on_the_fly(i, dim):
    """Column i (1-based) of the test operator (test_utils.f90:37-51 via expensive_function_1, :71-92)."""
    return _dv.compute_matrix_on_the_fly(_dv.OP_TEST_MTX, i, dim)


def compute_stx_on_the_fly(i, dim):
    """Column i of the test overlap operator (test_utils.f90:54-68 via expensive_function_2, :94-116)."""
    return _dv.compute_matrix_on_the_fly(_dv.OP_TEST_STX, i, dim)


def apply_mtx_to_vect(input_vect):
    """apply_mtx_to_vect (test_utils.f90:11-21): free_matmul(compute_matrix_on_the_fly, input_vect)."""
    return _dv.free_matmul(_dv.OP_TEST_MTX, input_vect)


def apply_stx_to_vect(input_vect):
    """apply_stx_to_vect (test_utils.f90:23-33)."""
    return _dv.free_matmul(_dv.OP_TEST_STX, input_vect)
