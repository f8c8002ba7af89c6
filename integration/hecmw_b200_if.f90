!-------------------------------------------------------------------------------
! ISO_C_BINDING interface of libfistr_b200.so (include/fistr_b200.h) for FrontISTR / HEC-MW.
!
! Drop this file into hecmw1/src/solver/b200/ and `use hecmw_b200_if` at the three call sites described in
! INTEGRATION.md (hecmw_solve, fstr_StiffMatrix/fstr_AddBC, fistr_main).  kint = 4 bytes = c_int and
! kreal = 8 bytes = c_double in HEC-MW (hecmw1/src/common/hecmw_util_f.F90:15-16), so hecMAT%D, hecMAT%indexL, ...
! are passed as they are; the 1-based contents of the index arrays are what the library expects.
!
! NOTE: no Fortran compiler exists in the image this repository is built in; this file is reviewed by eye and
! mirrored one-to-one by the ctypes declarations the tests use (frontistr_b200/hecmw.py).
!-------------------------------------------------------------------------------
module hecmw_b200_if
  use iso_c_binding
  implicit none

  integer(c_int), parameter :: FISTR_B200_OK = 0
  integer(c_int), parameter :: FISTR_B200_ERR_ARG = -1, FISTR_B200_ERR_CUDA = -2, FISTR_B200_ERR_COMM = -3
  integer(c_int), parameter :: FISTR_B200_ERR_STATE = -4, FISTR_B200_ERR_CONNECTIVITY = -5

  interface
    ! ---- life cycle -----------------------------------------------------------------------------
    integer(c_int) function fistr_b200_init(device_ordinal) bind(C, name="fistr_b200_init")
      import :: c_int
      integer(c_int), value :: device_ordinal          ! -1: LOCAL_RANK modulo the device count
    end function
    integer(c_int) function fistr_b200_finalize() bind(C, name="fistr_b200_finalize")
      import :: c_int
    end function
    integer(c_int) function fistr_b200_set_option(key, val) bind(C, name="fistr_b200_set_option")
      import :: c_int, c_char, c_int64_t
      character(kind=c_char) :: key(*)                  ! NUL-terminated
      integer(c_int64_t), value :: val
    end function

    ! ---- communicator (replaces MPI under hecmw_update_3_R / hecmw_allreduce_R1) -------------------
    integer(c_int) function fistr_b200_comm_get_unique_id(id128) bind(C, name="fistr_b200_comm_get_unique_id")
      import :: c_int, c_char
      character(kind=c_char) :: id128(128)
    end function
    integer(c_int) function fistr_b200_comm_init(id128, my_rank, petot) bind(C, name="fistr_b200_comm_init")
      import :: c_int, c_char
      character(kind=c_char) :: id128(128)
      integer(c_int), value :: my_rank, petot
    end function
    integer(c_int) function fistr_b200_comm_finalize() bind(C, name="fistr_b200_comm_finalize")
      import :: c_int
    end function

    ! ---- hecmw_mat_con ------------------------------------------------------------------------------
    integer(c_int) function fistr_b200_mat_con_count(n_node, nn_internal, n_elem_type, elem_type_index, &
        elem_type_item, elem_node_index, elem_node_item, N, NP, NPL, NPU) bind(C, name="fistr_b200_mat_con_count")
      import :: c_int
      integer(c_int), value :: n_node, nn_internal, n_elem_type
      integer(c_int) :: elem_type_index(0:*), elem_type_item(*), elem_node_index(0:*), elem_node_item(*)
      integer(c_int) :: N, NP, NPL, NPU
    end function
    integer(c_int) function fistr_b200_mat_con_fill(indexL, itemL, indexU, itemU) bind(C, name="fistr_b200_mat_con_fill")
      import :: c_int
      integer(c_int) :: indexL(0:*), itemL(*), indexU(0:*), itemU(*)
    end function

    ! ---- residency ----------------------------------------------------------------------------------
    integer(c_int) function fistr_b200_mat_structure(N, NP, NPL, NPU, indexL, itemL, indexU, itemU, n_neighbor_pe, &
        neighbor_pe, import_index, import_item, export_index, export_item) bind(C, name="fistr_b200_mat_structure")
      import :: c_int
      integer(c_int), value :: N, NP, NPL, NPU, n_neighbor_pe
      integer(c_int) :: indexL(0:*), itemL(*), indexU(0:*), itemU(*)
      integer(c_int) :: neighbor_pe(*), import_index(0:*), import_item(*), export_index(0:*), export_item(*)
    end function
    integer(c_int) function fistr_b200_mat_upload(D, AL, AU, B) bind(C, name="fistr_b200_mat_upload")
      import :: c_int, c_double
      real(c_double) :: D(*), AL(*), AU(*), B(*)
    end function
    integer(c_int) function fistr_b200_mat_download(D, AL, AU, B, X) bind(C, name="fistr_b200_mat_download")
      import :: c_int, c_double
      real(c_double) :: D(*), AL(*), AU(*), B(*), X(*)
    end function
    integer(c_int) function fistr_b200_host_register(ptr, bytes) bind(C, name="fistr_b200_host_register")
      import :: c_int, c_ptr, c_int64_t
      type(c_ptr), value :: ptr                         ! c_loc(hecMAT%AL(1)) ...
      integer(c_int64_t), value :: bytes
    end function

    ! ---- fstr_StiffMatrix + hecmw_mat_ass_elem, hecmw_mat_ass_bc ---------------------------------------
    integer(c_int) function fistr_b200_stiffmatrix_3d(n_node, node, n_elem, n_elem_type, elem_type_index, &
        elem_type_item, elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu, &
        D, AL, AU, time_ms) bind(C, name="fistr_b200_stiffmatrix_3d")
      import :: c_int, c_double
      integer(c_int), value :: n_node, n_elem, n_elem_type, n_sect
      real(c_double) :: node(*), sect_E(*), sect_nu(*), D(*), AL(*), AU(*), time_ms(2)
      integer(c_int) :: elem_type_index(0:*), elem_type_item(*), elem_node_index(0:*), elem_node_item(*)
      integer(c_int) :: section_ID(*), sect_elemopt361(*)
    end function
    ! ---- fstr_UpdateNewton (nlgeom_flag == INFINITE) + fstr_NodalStress3D: strain / stress recovery -------------
    ! Arguments that may be absent are C pointers by value: pass c_loc(array) or c_null_ptr
    ! (sect_alpha / temperature / temperature0: no !TEMPERATURE; QFORCE and any output of nodal_stress_3d: not wanted).
    integer(c_int) function fistr_b200_update_3d(n_node, node, n_elem, n_elem_type, elem_type_index, &
        elem_type_item, elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu, &
        sect_alpha, unode, dunode, temperature, temperature0, ref_temp, QFORCE) bind(C, name="fistr_b200_update_3d")
      import :: c_int, c_double, c_ptr
      integer(c_int), value :: n_node, n_elem, n_elem_type, n_sect
      real(c_double) :: node(*), sect_E(*), sect_nu(*), unode(*), dunode(*)
      integer(c_int) :: elem_type_index(0:*), elem_type_item(*), elem_node_index(0:*), elem_node_item(*)
      integer(c_int) :: section_ID(*), sect_elemopt361(*)
      type(c_ptr), value :: sect_alpha, temperature, temperature0, QFORCE
      real(c_double), value :: ref_temp
    end function
    integer(c_int) function fistr_b200_nodal_stress_3d(STRAIN, STRESS, MISES, ESTRAIN, ESTRESS, EMISES) &
        bind(C, name="fistr_b200_nodal_stress_3d")
      import :: c_int, c_ptr
      type(c_ptr), value :: STRAIN, STRESS, MISES, ESTRAIN, ESTRESS, EMISES   ! fstrSOLID%STRAIN(6*n_node) ... EMISES(n_elem)
    end function

    ! ---- geometrically nonlinear static (total Lagrangian): fstr_StiffMatrix / fstr_UpdateNewton under NLGEOM ---
    integer(c_int) function fistr_b200_nl_state_reset(n_elem) bind(C, name="fistr_b200_nl_state_reset")
      import :: c_int
      integer(c_int), value :: n_elem
    end function
    integer(c_int) function fistr_b200_nl_state_download(strain, stress) bind(C, name="fistr_b200_nl_state_download")
      import :: c_int, c_ptr
      type(c_ptr), value :: strain, stress          ! [n_elem][8][6] each, or c_null_ptr
    end function
    integer(c_int) function fistr_b200_nl_stiffmatrix_3d(n_node, node, n_elem, n_elem_type, elem_type_index, &
        elem_type_item, elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu, &
        sect_mtype, sect_hyper, u, D, AL, AU, time_ms) bind(C, name="fistr_b200_nl_stiffmatrix_3d")
      import :: c_int, c_double, c_ptr
      integer(c_int), value :: n_node, n_elem, n_elem_type, n_sect
      real(c_double) :: node(*), sect_E(*), sect_nu(*), u(*), time_ms(2)
      integer(c_int) :: elem_type_index(0:*), elem_type_item(*), elem_node_index(0:*), elem_node_item(*)
      integer(c_int) :: section_ID(*), sect_elemopt361(*)
      type(c_ptr), value :: sect_mtype, sect_hyper, D, AL, AU   ! c_null_ptr: all elastic / matrix stays on the device
    end function
    integer(c_int) function fistr_b200_nl_update_3d(n_node, node, n_elem, n_elem_type, elem_type_index, &
        elem_type_item, elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu, &
        sect_mtype, sect_hyper, unode, dunode, QFORCE) bind(C, name="fistr_b200_nl_update_3d")
      import :: c_int, c_double, c_ptr
      integer(c_int), value :: n_node, n_elem, n_elem_type, n_sect
      real(c_double) :: node(*), sect_E(*), sect_nu(*), unode(*), dunode(*), QFORCE(*)
      integer(c_int) :: elem_type_index(0:*), elem_type_item(*), elem_node_index(0:*), elem_node_item(*)
      integer(c_int) :: section_ID(*), sect_elemopt361(*)
      type(c_ptr), value :: sect_mtype, sect_hyper
    end function

    integer(c_int) function fistr_b200_mat_ass_bc(nbc, inode, idof, rhs) bind(C, name="fistr_b200_mat_ass_bc")
      import :: c_int, c_double
      integer(c_int), value :: nbc
      integer(c_int) :: inode(*), idof(*)
      real(c_double) :: rhs(*)
    end function

    ! ---- hecmw_solve ---------------------------------------------------------------------------------
    integer(c_int) function fistr_b200_solve_33(N, NP, NPL, NPU, D, AL, AU, indexL, itemL, indexU, itemU, &
        B, X, Iarray, Rarray, n_neighbor_pe, neighbor_pe, import_index, import_item, export_index, &
        export_item, my_rank, petot, iter_out, resid_out, time_out) bind(C, name="fistr_b200_solve_33")
      import :: c_int, c_double
      integer(c_int), value :: N, NP, NPL, NPU, n_neighbor_pe, my_rank, petot
      real(c_double)  :: D(*), AL(*), AU(*), B(*), X(*), Rarray(100), time_out(5)
      integer(c_int)  :: indexL(0:*), itemL(*), indexU(0:*), itemU(*), Iarray(100)
      integer(c_int)  :: neighbor_pe(*), import_index(0:*), import_item(*), export_index(0:*), export_item(*)
      integer(c_int)  :: iter_out
      real(c_double)  :: resid_out
    end function

    ! ---- hecmw_matvec_33, hecmw_precond_33_*, hecmw_update_3_R -------------------------------------------
    integer(c_int) function fistr_b200_matvec_33(X, Y) bind(C, name="fistr_b200_matvec_33")
      import :: c_int, c_double
      real(c_double) :: X(*), Y(*)
    end function
    integer(c_int) function fistr_b200_precond_33_setup(precond, sigma_diag, ncolor_in) bind(C, name="fistr_b200_precond_33_setup")
      import :: c_int, c_double
      integer(c_int), value :: precond, ncolor_in
      real(c_double), value :: sigma_diag
    end function
    integer(c_int) function fistr_b200_precond_33_apply(ZP) bind(C, name="fistr_b200_precond_33_apply")
      import :: c_int, c_double
      real(c_double) :: ZP(*)
    end function
    integer(c_int) function fistr_b200_precond_33_clear() bind(C, name="fistr_b200_precond_33_clear")
      import :: c_int
    end function
    integer(c_int) function fistr_b200_update_3_R(X) bind(C, name="fistr_b200_update_3_R")
      import :: c_int, c_double
      real(c_double) :: X(*)
    end function
  end interface

contains

  !> what `hecmw_solve` calls instead of `hecmw_solve_iterative` for NDOF=3, METHOD=CG (INTEGRATION.md 3a)
  subroutine hecmw_b200_solve(N, NP, NPL, NPU, D, AL, AU, indexL, itemL, indexU, itemU, B, X, Iarray, Rarray, &
      n_neighbor_pe, neighbor_pe, import_index, import_item, export_index, export_item, my_rank, petot, ierr)
    integer(c_int), intent(in) :: N, NP, NPL, NPU, n_neighbor_pe, my_rank, petot
    real(c_double)  :: D(*), AL(*), AU(*), B(*), X(*), Rarray(100)
    integer(c_int)  :: indexL(0:*), itemL(*), indexU(0:*), itemU(*), Iarray(100)
    integer(c_int)  :: neighbor_pe(*), import_index(0:*), import_item(*), export_index(0:*), export_item(*)
    integer(c_int), intent(out) :: ierr
    integer(c_int) :: iter
    real(c_double) :: resid, times(5)
    ierr = fistr_b200_solve_33(N, NP, NPL, NPU, D, AL, AU, indexL, itemL, indexU, itemU, B, X, Iarray, Rarray, &
      n_neighbor_pe, neighbor_pe, import_index, import_item, export_index, export_item, my_rank, petot, &
      iter, resid, times)
    ! ierr: 0 ok | 2002, 3001-3003 = HEC-MW warnings (hecmw_solve_error.f90) | 1001, 2001 or < 0 -> caller aborts
  end subroutine hecmw_b200_solve

end module hecmw_b200_if
