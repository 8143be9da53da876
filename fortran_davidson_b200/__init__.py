"""fortran_davidson_b200 -- B200-native block Davidson eigensolver behind the `davidson` module API
of NLESC-JCER/Fortran_Davidson.  Python here is only the host-side mirror of the reference's
interface (same names, argument meaning and error behaviour) over the C ABI of
libdavidson_b200.so; all numerics run in hand-written sm_100a CUDA kernels."""
from . import array_utils, lapack_wrapper  # noqa: F401
from .array_utils import generate_diagonal_dominant  # noqa: F401  (README.md:26 imports it from `davidson`)
from .davidson import (DavidsonSolver, eigensolver, free_matmul, generalized_eigensolver,  # noqa: F401
                       OP_BENCHMARK_MTX, OP_IDENTITY, OP_TEST_MTX, OP_TEST_STX)
from ._lib import DavidsonError, lib  # noqa: F401
