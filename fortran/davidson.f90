!> Drop-in replacement of the reference's `davidson`, `array_utils`, `lapack_wrapper` and
!> `numeric_kinds` modules (NLESC-JCER/Fortran_Davidson, src/davidson.f90, src/array_utils.f90,
!> src/lapack_wrapper.f90, src/numeric_kinds.f90): the same public names and argument lists, every
!> computation forwarded through `iso_c_binding` to the C ABI of libdavidson_b200.so
!> (include/davidson_b200.h), whose kernels run on the B200.
!>
!> Build (where a Fortran compiler exists -- there is none in the image this repository is developed in,
!> so this file is shipped as source and its C side is what the test-suite exercises):
!>     gfortran -c fortran/davidson.f90
!>     gfortran main.f90 davidson.o -L<repo>/fortran_davidson_b200 -ldavidson_b200 -Wl,-rpath,<...>
!>
!> Error convention: the reference prints and `error stop`s when LAPACK fails
!> (lapack_wrapper.f90:395-408); here any non-zero status of the C ABI does the same with
!> dav_last_error()'s text.  Non-convergence stays a warning with iters = max_iterations+1
!> (davidson.f90:232-235), printed by the library.
!>
!> Threading: the matrix-free path keeps the two user procedures in module variables while a solve
!> is running (the C callback has to find them), so one matrix-free solve at a time per process.

module numeric_kinds
  use iso_c_binding, only: c_double, c_float, c_int32_t, c_int64_t
  implicit none
  integer, parameter :: dp = c_double
  integer, parameter :: sp = c_float
  integer, parameter :: qp = selected_real_kind(2 * precision(1.0_dp))
  integer, parameter :: i4b = c_int32_t
  integer, parameter :: i8b = c_int64_t
end module numeric_kinds


!> raw bind(C) interfaces of include/davidson_b200.h
module davidson_b200_c
  use iso_c_binding
  implicit none

  abstract interface
     subroutine dav_gemv_fn(x, y, n, b, ctx) bind(C)
       import :: c_ptr, c_int64_t
       type(c_ptr), value :: x, y
       integer(c_int64_t), value :: n, b
       type(c_ptr), value :: ctx
     end subroutine dav_gemv_fn
  end interface

  interface
     function dav_last_error() bind(C, name="dav_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function dav_last_error

     function dav_generalized_eigensolver_dense(n, matrix, lda, second_matrix, ldb, lowest, method, &
          max_iterations, tolerance, max_dim_sub, eigenvalues, eigenvectors, ldv, iters) &
          bind(C, name="dav_generalized_eigensolver_dense") result(ierr)
       import :: c_ptr, c_int, c_int64_t, c_double, c_char
       integer(c_int64_t), value :: n, lda, ldb, ldv
       type(c_ptr), value :: matrix, second_matrix          ! second_matrix = c_null_ptr when absent
       integer(c_int), value :: lowest, max_iterations, max_dim_sub
       character(kind=c_char), dimension(*), intent(in) :: method
       real(c_double), value :: tolerance
       real(c_double), intent(out) :: eigenvalues(*)
       real(c_double), intent(out) :: eigenvectors(ldv, *)
       integer(c_int), intent(out) :: iters
       integer(c_int) :: ierr
     end function dav_generalized_eigensolver_dense

     function dav_generalized_eigensolver_free(n, fun_matrix_gemv, ctx_matrix, fun_second_matrix_gemv, ctx_second, &
          diag_matrix, diag_second_matrix, lowest, method, max_iterations, tolerance, max_dim_sub, &
          eigenvalues, ritz_vectors, ldv, iters) bind(C, name="dav_generalized_eigensolver_free") result(ierr)
       import :: c_ptr, c_funptr, c_int, c_int64_t, c_double, c_char
       integer(c_int64_t), value :: n, ldv
       type(c_funptr), value :: fun_matrix_gemv, fun_second_matrix_gemv
       type(c_ptr), value :: ctx_matrix, ctx_second, diag_matrix, diag_second_matrix
       integer(c_int), value :: lowest, max_iterations, max_dim_sub
       character(kind=c_char), dimension(*), intent(in) :: method
       real(c_double), value :: tolerance
       real(c_double), intent(out) :: eigenvalues(*)
       real(c_double), intent(out) :: ritz_vectors(ldv, *)
       integer(c_int), intent(inout) :: iters
       integer(c_int) :: ierr
     end function dav_generalized_eigensolver_free

     function dav_generate_diagonal_dominant(m, sparsity, diag_val, seed, arr, ld) &
          bind(C, name="dav_generate_diagonal_dominant") result(ierr)
       import :: c_ptr, c_int, c_int64_t, c_double
       integer(c_int64_t), value :: m, ld, seed
       real(c_double), value :: sparsity
       type(c_ptr), value :: diag_val                         ! c_null_ptr when absent
       real(c_double), intent(out) :: arr(ld, *)
       integer(c_int) :: ierr
     end function dav_generate_diagonal_dominant

     function dav_generate_preconditioner(n, diag, dim_sub, precond, ld) &
          bind(C, name="dav_generate_preconditioner") result(ierr)
       import :: c_int, c_int64_t, c_double
       integer(c_int64_t), value :: n, ld
       real(c_double), intent(in) :: diag(*)
       integer(c_int), value :: dim_sub
       real(c_double), intent(out) :: precond(ld, *)
       integer(c_int) :: ierr
     end function dav_generate_preconditioner

     function dav_norm(n, vector, res) bind(C, name="dav_norm") result(ierr)
       import :: c_int, c_int64_t, c_double
       integer(c_int64_t), value :: n
       real(c_double), intent(in) :: vector(*)
       real(c_double), intent(out) :: res
       integer(c_int) :: ierr
     end function dav_norm

     pure function dav_norm_value(n, vector) bind(C, name="dav_norm_value") result(res)
       import :: c_int64_t, c_double
       integer(c_int64_t), value :: n
       real(c_double), intent(in) :: vector(*)
       real(c_double) :: res
     end function dav_norm_value

     function dav_lapack_generalized_eigensolver(dim, mtx, stx, eigenvalues, eigenvectors) &
          bind(C, name="dav_lapack_generalized_eigensolver") result(ierr)
       import :: c_ptr, c_int, c_double
       integer(c_int), value :: dim
       real(c_double), intent(in) :: mtx(dim, *)
       type(c_ptr), value :: stx                              ! c_null_ptr when absent
       real(c_double), intent(out) :: eigenvalues(*), eigenvectors(dim, *)
       integer(c_int) :: ierr
     end function dav_lapack_generalized_eigensolver

     function dav_lapack_generalized_eigensolver_lowest(dim, mtx, stx, lowest, eigenvalues, eigenvectors) &
          bind(C, name="dav_lapack_generalized_eigensolver_lowest") result(ierr)
       import :: c_int, c_double
       integer(c_int), value :: dim, lowest
       real(c_double), intent(in) :: mtx(dim, *), stx(dim, *)
       real(c_double), intent(out) :: eigenvalues(*), eigenvectors(dim, *)
       integer(c_int) :: ierr
     end function dav_lapack_generalized_eigensolver_lowest

     function dav_lapack_qr(m, n, basis, ld) bind(C, name="dav_lapack_qr") result(ierr)
       import :: c_int, c_int64_t, c_double
       integer(c_int64_t), value :: m, ld
       integer(c_int), value :: n
       real(c_double), intent(inout) :: basis(ld, *)
       integer(c_int) :: ierr
     end function dav_lapack_qr

     function dav_lapack_solver(n, arr, brr) bind(C, name="dav_lapack_solver") result(ierr)
       import :: c_int, c_double
       integer(c_int), value :: n
       real(c_double), intent(in) :: arr(n, *)
       real(c_double), intent(inout) :: brr(*)
       integer(c_int) :: ierr
     end function dav_lapack_solver

     function dav_lapack_matmul(transA, transB, rows_a, cols_a, arr, rows_b, cols_b, brr, alpha, mtx) &
          bind(C, name="dav_lapack_matmul") result(ierr)
       import :: c_int, c_int64_t, c_double, c_char
       character(kind=c_char), value :: transA, transB
       integer(c_int64_t), value :: rows_a, cols_a, rows_b, cols_b
       real(c_double), intent(in) :: arr(rows_a, *), brr(rows_b, *)
       real(c_double), value :: alpha
       real(c_double), intent(out) :: mtx(*)
       integer(c_int) :: ierr
     end function dav_lapack_matmul

     function dav_lapack_matrix_vector(transA, m, n, mtx, vector, alpha, rs) &
          bind(C, name="dav_lapack_matrix_vector") result(ierr)
       import :: c_int, c_int64_t, c_double, c_char
       character(kind=c_char), value :: transA
       integer(c_int64_t), value :: m, n
       real(c_double), intent(in) :: mtx(m, *), vector(*)
       real(c_double), value :: alpha
       real(c_double), intent(out) :: rs(*)
       integer(c_int) :: ierr
     end function dav_lapack_matrix_vector

     function dav_lapack_sort(id, n, vector, keys) bind(C, name="dav_lapack_sort") result(ierr)
       import :: c_int, c_int32_t, c_int64_t, c_double, c_char
       character(kind=c_char), value :: id
       integer(c_int64_t), value :: n
       real(c_double), intent(inout) :: vector(*)
       integer(c_int32_t), intent(out) :: keys(*)
       integer(c_int) :: ierr
     end function dav_lapack_sort
  end interface

contains

  !> print dav_last_error() and `error stop`, like check_lapack_call (lapack_wrapper.f90:395-408)
  subroutine check_status(ierr, name)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: name
    character(kind=c_char), pointer :: chars(:)
    type(c_ptr) :: msg
    integer :: i
    if (ierr == 0) return
    print *, "call to subroutine: ", name, " has failed!"
    print *, "info: ", ierr
    msg = dav_last_error()
    if (c_associated(msg)) then
       call c_f_pointer(msg, chars, [512])
       i = 1
       do while (i <= 512)
          if (chars(i) == c_null_char) exit
          i = i + 1
       end do
       print *, chars(1:i - 1)
    end if
    error stop
  end subroutine check_status

end module davidson_b200_c


module lapack_wrapper
  !> same public list as lapack_wrapper.f90:9-10
  use iso_c_binding
  use numeric_kinds, only: dp
  use davidson_b200_c
  implicit none
  private
  public :: lapack_generalized_eigensolver, lapack_generalized_eigensolver_lowest, &
       lapack_matmul, lapack_matrix_vector, lapack_qr, lapack_solver, lapack_sort

contains

  subroutine lapack_generalized_eigensolver(mtx, eigenvalues, eigenvectors, stx)
    !> lapack_wrapper.f90:14-91
    real(dp), dimension(:, :), intent(in) :: mtx
    real(dp), dimension(:, :), intent(in), optional, target :: stx
    real(dp), dimension(size(mtx, 1)), intent(inout) :: eigenvalues
    real(dp), dimension(size(mtx, 1), size(mtx, 2)), intent(inout) :: eigenvectors
    real(dp), dimension(:, :), allocatable, target :: stx_copy
    real(dp), dimension(:, :), allocatable :: mtx_copy
    type(c_ptr) :: pstx
    mtx_copy = mtx                                   ! contiguous copy (assumed-shape dummy)
    pstx = c_null_ptr
    if (present(stx)) then
       stx_copy = stx
       pstx = c_loc(stx_copy)
    end if
    call check_status(dav_lapack_generalized_eigensolver(int(size(mtx, 1), c_int), mtx_copy, pstx, &
         eigenvalues, eigenvectors), "DSYGV")
  end subroutine lapack_generalized_eigensolver

  subroutine lapack_generalized_eigensolver_lowest(mtx, stx, eigenvalues, eigenvectors, lowest)
    !> lapack_wrapper.f90:93-174
    integer :: lowest
    real(dp), dimension(:, :), intent(in) :: mtx, stx
    real(dp), dimension(lowest), intent(out) :: eigenvalues
    real(dp), dimension(size(mtx, 1), lowest), intent(out) :: eigenvectors
    real(dp), dimension(:, :), allocatable :: mtx_copy, stx_copy
    mtx_copy = mtx
    stx_copy = stx
    call check_status(dav_lapack_generalized_eigensolver_lowest(int(size(mtx, 1), c_int), mtx_copy, stx_copy, &
         int(lowest, c_int), eigenvalues, eigenvectors), "DSYGVX")
  end subroutine lapack_generalized_eigensolver_lowest

  subroutine lapack_qr(basis)
    !> lapack_wrapper.f90:176-236
    real(dp), dimension(:, :), intent(inout) :: basis
    real(dp), dimension(:, :), allocatable :: work
    work = basis
    call check_status(dav_lapack_qr(int(size(basis, 1), c_int64_t), int(size(basis, 2), c_int), work, &
         int(size(basis, 1), c_int64_t)), "DGEQRF")
    basis = work
  end subroutine lapack_qr

  subroutine lapack_solver(arr, brr)
    !> lapack_wrapper.f90:238-277
    real(dp), dimension(:, :), intent(inout) :: arr, brr
    real(dp), dimension(:, :), allocatable :: a
    real(dp), dimension(:), allocatable :: b
    a = arr
    b = brr(:, 1)
    call check_status(dav_lapack_solver(int(size(arr, 1), c_int), a, b), "DSYSV")
    brr(:, 1) = b
  end subroutine lapack_solver

  function lapack_matmul(transA, transB, arr, brr, alpha) result(mtx)
    !> lapack_wrapper.f90:279-328
    character(len=1), intent(in) :: transA, transB
    real(dp), dimension(:, :), intent(in) :: arr, brr
    real(dp), optional, intent(in) :: alpha
    real(dp), dimension(:, :), allocatable :: mtx, a, b
    real(dp) :: x
    integer :: m, n
    x = 1.d0
    if (present(alpha)) x = alpha
    m = merge(size(arr, 2), size(arr, 1), transA == 'T')
    n = merge(size(brr, 1), size(brr, 2), transB == 'T')
    allocate(mtx(m, n))
    a = arr
    b = brr
    call check_status(dav_lapack_matmul(transA, transB, int(size(arr, 1), c_int64_t), int(size(arr, 2), c_int64_t), a, &
         int(size(brr, 1), c_int64_t), int(size(brr, 2), c_int64_t), b, x, mtx), "DGEMM")
  end function lapack_matmul

  function lapack_matrix_vector(transA, mtx, vector, alpha) result(rs)
    !> lapack_wrapper.f90:330-364
    character(len=1), intent(in) :: transA
    real(dp), dimension(:, :), intent(in) :: mtx
    real(dp), dimension(:), intent(in) :: vector
    real(dp), optional, intent(in) :: alpha
    real(dp), dimension(:), allocatable :: rs, v
    real(dp), dimension(:, :), allocatable :: a
    real(dp) :: scalar
    scalar = 1.d0
    if (present(alpha)) scalar = alpha
    ! op(mtx) * vector has size(mtx, 2) entries for transA = 'T' (the reference allocates size(mtx, 1) in both
    ! cases, lapack_wrapper.f90:349, which only works for square matrices)
    if (transA == 'T' .or. transA == 't') then
       allocate(rs(size(mtx, 2)))
    else
       allocate(rs(size(mtx, 1)))
    end if
    rs = 0.d0
    a = mtx
    v = vector
    call check_status(dav_lapack_matrix_vector(transA, int(size(mtx, 1), c_int64_t), int(size(mtx, 2), c_int64_t), a, v, &
         scalar, rs), "DGEMV")
  end function lapack_matrix_vector

  function lapack_sort(id, vector) result(keys)
    !> lapack_wrapper.f90:367-392 (sorts `vector` in place, returns the rank of every original element)
    real(dp), dimension(:), intent(inout) :: vector
    character(len=1), intent(in) :: id
    integer, dimension(size(vector)) :: keys
    real(dp), dimension(:), allocatable :: v
    integer(c_int32_t), dimension(:), allocatable :: k32
    v = vector
    allocate(k32(size(vector)))
    call check_status(dav_lapack_sort(id, int(size(vector), c_int64_t), v, k32), "DLASRT")
    vector = v
    keys = int(k32)
  end function lapack_sort

end module lapack_wrapper


module array_utils
  !> same public list as array_utils.f90:11-12
  use iso_c_binding
  use numeric_kinds, only: dp
  use davidson_b200_c
  implicit none
  private
  public :: concatenate, diagonal, eye, generate_diagonal_dominant, norm, generate_preconditioner

contains

  pure function eye(m, n, alpha)
    !> array_utils.f90:16-44
    integer, intent(in) :: n, m
    real(dp), dimension(m, n) :: eye
    real(dp), intent(in), optional :: alpha
    integer :: i
    real(dp) :: x
    x = 1.d0
    if (present(alpha)) x = alpha
    eye = 0.d0
    do i = 1, min(m, n)
       eye(i, i) = x
    end do
  end function eye

  pure function norm(vector)
    !> array_utils.f90:46-53 -- `pure` like the reference, so that callers may use it in pure / elemental contexts.
    !> dav_norm_value has no side effect a Fortran program can observe and returns NaN instead of an error code.
    real(dp), dimension(:), intent(in) :: vector
    real(dp) :: norm
    norm = dav_norm_value(int(size(vector), c_int64_t), vector)
  end function norm

  subroutine concatenate(arr, brr)
    !> array_utils.f90:55-84
    real(dp), dimension(:, :), intent(inout), allocatable :: arr
    real(dp), dimension(:, :), intent(in) :: brr
    real(dp), dimension(:, :), allocatable :: tmp_array
    integer :: dim_cols
    dim_cols = size(arr, 2)
    allocate(tmp_array(size(arr, 1), dim_cols + size(brr, 2)))
    tmp_array(:, :dim_cols) = arr
    tmp_array(:, dim_cols + 1:) = brr
    call move_alloc(tmp_array, arr)
  end subroutine concatenate

  function generate_diagonal_dominant(m, sparsity, diag_val) result(arr)
    !> array_utils.f90:86-113.  The reference draws from the compiler's unseeded PRNG; here the entries come
    !> from the library's counter-based stream (seed 0), generated on the GPU.
    integer, intent(in) :: m
    real(dp), optional, target :: diag_val
    real(dp) :: sparsity
    real(dp), dimension(m, m) :: arr
    type(c_ptr) :: pdiag
    pdiag = c_null_ptr
    if (present(diag_val)) pdiag = c_loc(diag_val)
    call check_status(dav_generate_diagonal_dominant(int(m, c_int64_t), sparsity, pdiag, 0_c_int64_t, arr, &
         int(m, c_int64_t)), "generate_diagonal_dominant")
  end function generate_diagonal_dominant

  function diagonal(matrix)
    !> array_utils.f90:115-134
    real(dp), dimension(:, :), intent(in) :: matrix
    real(dp), dimension(size(matrix, 1)) :: diagonal
    integer :: i
    do i = 1, size(matrix, 1)
       diagonal(i) = matrix(i, i)
    end do
  end function diagonal

  function generate_preconditioner(diag, dim_sub) result(precond)
    !> array_utils.f90:136-160 (the reference also sorts `diag` in place as a side effect; kept)
    real(dp), dimension(:), intent(inout) :: diag
    integer, intent(in) :: dim_sub
    real(dp), dimension(size(diag), dim_sub) :: precond
    real(dp), dimension(:), allocatable :: d
    integer(c_int32_t), dimension(:), allocatable :: keys
    d = diag
    call check_status(dav_generate_preconditioner(int(size(diag), c_int64_t), d, int(dim_sub, c_int), precond, &
         int(size(diag), c_int64_t)), "generate_preconditioner")
    allocate(keys(size(diag)))
    call check_status(dav_lapack_sort('I', int(size(diag), c_int64_t), d, keys), "DLASRT")
    diag = d
  end function generate_preconditioner

end module array_utils


module davidson_dense
  !> generalized_eigensolver_dense (davidson.f90:51-246)
  use iso_c_binding
  use numeric_kinds, only: dp
  use davidson_b200_c
  implicit none
  private
  public :: generalized_eigensolver_dense

contains

  subroutine generalized_eigensolver_dense(matrix, eigenvalues, eigenvectors, lowest, method, max_iterations, &
       tolerance, iters, max_dim_sub, second_matrix)
    integer, intent(in) :: lowest
    real(dp), dimension(:, :), intent(in) :: matrix
    real(dp), dimension(:, :), intent(in), optional :: second_matrix
    real(dp), dimension(lowest), intent(out) :: eigenvalues
    real(dp), dimension(:, :), intent(out) :: eigenvectors
    integer, intent(in) :: max_iterations
    integer, intent(in), optional :: max_dim_sub
    real(dp), intent(in) :: tolerance
    character(len=*), intent(in) :: method
    integer, intent(out) :: iters

    real(dp), dimension(:, :), allocatable, target :: a, b
    real(dp), dimension(:, :), allocatable :: vec
    type(c_ptr) :: pb
    integer(c_int) :: c_iters, mds
    integer(c_int64_t) :: n

    n = size(matrix, 1)
    a = matrix                                        ! contiguous, caller's array is never modified
    pb = c_null_ptr
    if (present(second_matrix)) then
       b = second_matrix
       pb = c_loc(b)
    end if
    mds = 0                                           ! 0 = "not present" (default 10*lowest, davidson.f90:115-119)
    if (present(max_dim_sub)) mds = int(max_dim_sub, c_int)
    allocate(vec(n, lowest))
    call check_status(dav_generalized_eigensolver_dense(n, c_loc(a), n, pb, n, int(lowest, c_int), &
         trim(method) // c_null_char, int(max_iterations, c_int), tolerance, mds, eigenvalues, vec, n, c_iters), &
         "generalized_eigensolver_dense")
    eigenvectors(:, :lowest) = vec
    iters = int(c_iters)
  end subroutine generalized_eigensolver_dense

end module davidson_dense


module davidson_free
  !> generalized_eigensolver_free and free_matmul (davidson.f90:277-460, :526-569)
  use iso_c_binding
  use numeric_kinds, only: dp
  use davidson_b200_c
  implicit none
  private
  public :: generalized_eigensolver_free, free_matmul

  abstract interface
     function gemv_like(input_vect) result(output_vect)
       use numeric_kinds, only: dp
       real(dp), dimension(:, :), intent(in) :: input_vect
       real(dp), dimension(size(input_vect, 1), size(input_vect, 2)) :: output_vect
     end function gemv_like
  end interface

  ! the two user procedures of the solve in flight (found by the bind(C) trampolines)
  procedure(gemv_like), pointer, save :: current_matrix_gemv => null()
  procedure(gemv_like), pointer, save :: current_second_matrix_gemv => null()

contains

  subroutine tramp_matrix(x, y, n, b, ctx) bind(C)
    type(c_ptr), value :: x, y
    integer(c_int64_t), value :: n, b
    type(c_ptr), value :: ctx
    real(dp), pointer :: xf(:, :), yf(:, :)
    call c_f_pointer(x, xf, [n, b])
    call c_f_pointer(y, yf, [n, b])
    yf = current_matrix_gemv(xf)
  end subroutine tramp_matrix

  subroutine tramp_second(x, y, n, b, ctx) bind(C)
    type(c_ptr), value :: x, y
    integer(c_int64_t), value :: n, b
    type(c_ptr), value :: ctx
    real(dp), pointer :: xf(:, :), yf(:, :)
    call c_f_pointer(x, xf, [n, b])
    call c_f_pointer(y, yf, [n, b])
    yf = current_second_matrix_gemv(xf)
  end subroutine tramp_second

  subroutine generalized_eigensolver_free(fun_matrix_gemv, eigenvalues, ritz_vectors, lowest, method, max_iterations, &
       tolerance, iters, max_dim_sub, fun_second_matrix_gemv)
    integer, intent(in) :: lowest
    real(dp), dimension(lowest), intent(out) :: eigenvalues
    real(dp), dimension(:, :), intent(out) :: ritz_vectors
    integer, intent(in) :: max_iterations
    integer, intent(in), optional :: max_dim_sub
    real(dp), intent(in) :: tolerance
    character(len=*), intent(in) :: method
    !> the reference declares intent(out) but leaves `iters` unassigned when the loop does not converge
    !> (davidson.f90:417); its value on entry is therefore handed through, which needs intent(inout)
    integer, intent(inout) :: iters
    procedure(gemv_like) :: fun_matrix_gemv, fun_second_matrix_gemv

    real(dp), dimension(:, :), allocatable :: vec
    integer(c_int) :: c_iters, mds
    integer(c_int64_t) :: n

    n = size(ritz_vectors, 1)                         ! davidson.f90:362
    current_matrix_gemv => fun_matrix_gemv
    current_second_matrix_gemv => fun_second_matrix_gemv
    mds = 0
    if (present(max_dim_sub)) mds = int(max_dim_sub, c_int)
    allocate(vec(n, lowest))
    c_iters = int(iters, c_int)                       ! left untouched when the loop does not converge (:417)
    ! diagonals are extracted by operator applications like extract_diagonal_free (davidson.f90:490-523)
    call check_status(dav_generalized_eigensolver_free(n, c_funloc(tramp_matrix), c_null_ptr, c_funloc(tramp_second), &
         c_null_ptr, c_null_ptr, c_null_ptr, int(lowest, c_int), trim(method) // c_null_char, &
         int(max_iterations, c_int), tolerance, mds, eigenvalues, vec, n, c_iters), "generalized_eigensolver_free")
    ritz_vectors(:, :lowest) = vec
    iters = int(c_iters)
    current_matrix_gemv => null()
    current_second_matrix_gemv => null()
  end subroutine generalized_eigensolver_free

  function free_matmul(fun, array) result(matrix)
    !> davidson.f90:526-569: matrix(i, j) = dot_product(fun(i, dim), array(:, j)).  Arbitrary user
    !> generators stay on the host; the built-in generators of benchmark_free.f90 / test_utils.f90 run on
    !> the device through dav_free_matmul / the DAV_OP_* operators of the handle API.
    real(dp), dimension(:, :), intent(in) :: array
    real(dp), dimension(size(array, 1), size(array, 2)) :: matrix
    interface
       function fun(i, dim) result(vec)
         use numeric_kinds, only: dp
         integer, intent(in) :: i
         integer, intent(in) :: dim
         real(dp), dimension(dim) :: vec
       end function fun
    end interface
    real(dp), dimension(size(array, 1)) :: vec
    integer :: dim1, dim2, i, j
    dim1 = size(array, 1)
    dim2 = size(array, 2)
    !$OMP PARALLEL DO PRIVATE(i, j, vec)
    do i = 1, dim1
       vec = fun(i, dim1)
       do j = 1, dim2
          matrix(i, j) = dot_product(vec, array(:, j))
       end do
    end do
    !$OMP END PARALLEL DO
  end function free_matmul

end module davidson_free


module davidson
  !> davidson.f90:586-627 plus the two names README.md:20,26 promises
  use numeric_kinds, only: dp
  use davidson_dense, only: generalized_eigensolver_dense
  use davidson_free, only: generalized_eigensolver_free, free_matmul
  use array_utils, only: generate_diagonal_dominant
  implicit none
  private
  public :: generalized_eigensolver, eigensolver, generate_diagonal_dominant, free_matmul

  interface generalized_eigensolver
     procedure generalized_eigensolver_dense
     procedure generalized_eigensolver_free
  end interface generalized_eigensolver

contains

  subroutine eigensolver(matrix, eigenvalues, eigenvectors, lowest, method, max_iterations, tolerance, iters, max_dim_sub)
    !> standard eigenvalue problem (README.md:20): generalized_eigensolver without second_matrix
    integer, intent(in) :: lowest
    real(dp), dimension(:, :), intent(in) :: matrix
    real(dp), dimension(lowest), intent(out) :: eigenvalues
    real(dp), dimension(:, :), intent(out) :: eigenvectors
    integer, intent(in) :: max_iterations
    integer, intent(in), optional :: max_dim_sub
    real(dp), intent(in) :: tolerance
    character(len=*), intent(in) :: method
    integer, intent(out) :: iters
    if (present(max_dim_sub)) then
       call generalized_eigensolver_dense(matrix, eigenvalues, eigenvectors, lowest, method, max_iterations, &
            tolerance, iters, max_dim_sub)
    else
       call generalized_eigensolver_dense(matrix, eigenvalues, eigenvectors, lowest, method, max_iterations, &
            tolerance, iters)
    end if
  end subroutine eigensolver

end module davidson
