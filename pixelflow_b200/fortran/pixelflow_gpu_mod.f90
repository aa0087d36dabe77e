! pixelflow_gpu_mod.f90 -- iso_c_binding interface to libpixelflow_gpu.so (include/pixelflow_gpu.h).
!
! Thin by design: argument marshalling only, no numerics.  A PixelFlow driver keeps its namelists
! (lib/global.f90), its porosity CSV reader and grid set-up (lib/grid.f90) and its output routines
! (lib/output.f90) and replaces the body of the time loop
! (src/omp_parallel/ibm_3d_uniform_omp_cpu.f90:81-132) by calls into this module -- see
! INTEGRATION.md for the exact hunk and pixelflow_b200/fortran/ibm3_uniform_gpu.f90 for a complete
! driver.  Build the reference library with -fdefault-real-8 (the GPU path is fp64) and link with
! -lpixelflow_gpu.
!
! NOTE: no Fortran compiler exists in the image this repository is developed in (SURVEY.md 0.7), so
! this file is delivered as source; the C++ twin driver (pixelflow_b200/driver) exercises the same
! C entry points and is what the tests run.
module pixelflow_gpu
  use iso_c_binding
  implicit none
  private

  integer(c_int), parameter, public :: PF_IBM2_UNIFORM = 0, PF_IBM2_BACKSTEP = 1, PF_IBM2_DRAG = 2, &
                                       PF_IBM3_UNIFORM = 3, PF_IBM3_AIRCOND = 4
  ! indices into pf_config%wall (+1: Fortran arrays are 1-based): top, bottom, east, west, south, north
  integer, parameter, public :: PF_TOP = 1, PF_BOTTOM = 2, PF_EAST = 3, PF_WEST = 4, PF_SOUTH = 5, PF_NORTH = 6

  ! mirrors `struct pf_config` field by field
  type, bind(C), public :: pf_config
    integer(c_int) :: struct_size
    integer(c_int) :: solver_case
    integer(c_int) :: m, n, l
    integer(c_int) :: host_ldx, host_ldy, host_is_slab
    real(c_double) :: dx, dy, dz, dt
    real(c_double) :: xnue, xlambda, density, thickness
    integer(c_int) :: nonslip
    integer(c_int) :: iter_max
    real(c_double) :: relux_factor
    real(c_double) :: inlet_velocity, outlet_pressure, AoA
    integer(c_int) :: wall(6)
    integer(c_int) :: device
    integer(c_int) :: rank, nranks
    type(c_ptr)    :: nccl_unique_id
    integer(c_int) :: sor_variant
    integer(c_int) :: use_graph
    integer(c_int) :: halo_transport
  end type pf_config

  public :: pf_config_init, pf_create, pf_destroy, pf_last_error, pf_set_porosity, pf_upload, pf_download
  public :: pf_step, pf_step_host, pf_initial_conditions, pf_copy_old, pf_divergence, pf_predictor
  public :: pf_build_poisson, pf_sor, pf_project, pf_boundary, pf_sync, pf_last_timing, pf_local_slab
  public :: pf_check, pf_error_message, pf_force_log_2d, pf_force_log_3d, pf_vtk_section_bytes, pf_vtk_section
  ! several GPUs from this one program (no launcher): INTEGRATION.md, "Several GPUs from one driver"
  public :: pf_ranks_launch, pf_ranks_rank, pf_ranks_count, pf_ranks_unique_id, pf_ranks_barrier, pf_ranks_finish
  public :: pf_gather
  ! enum pf_vtk_section: the sections of output_paraview_temp_3d / _2d in file order
  integer(c_int), parameter, public :: PF_VTK_POINTS = 0, PF_VTK_VELOCITY = 1, PF_VTK_VELOCITY_IN_FLUID = 2, &
       PF_VTK_DIMLESS_V = 3, PF_VTK_POROSITY = 4, PF_VTK_PRESSURE = 5, PF_VTK_DIVERGENT = 6, PF_VTK_ABS_DIMLESS_V = 7

  interface
    subroutine pf_config_init(cfg) bind(C, name="pf_config_init")
      import :: pf_config
      type(pf_config), intent(out) :: cfg
    end subroutine
    integer(c_int) function pf_create(handle, cfg) bind(C, name="pf_create")
      import :: c_int, c_ptr, pf_config
      type(c_ptr), intent(out) :: handle
      type(pf_config), intent(in) :: cfg
    end function
    subroutine pf_destroy(handle) bind(C, name="pf_destroy")
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine
    type(c_ptr) function pf_last_error(handle) bind(C, name="pf_last_error")
      import :: c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_local_slab(handle, k_first, k_count) bind(C, name="pf_local_slab")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), intent(out) :: k_first, k_count
    end function
    ! arrays are passed as the first element of the Fortran array: real(8), dimension(0:md,0:nd,0:ld)
    integer(c_int) function pf_set_porosity(handle, porosity) bind(C, name="pf_set_porosity")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: porosity(*)
    end function
    integer(c_int) function pf_upload(handle, u, v, w, p) bind(C, name="pf_upload")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: u(*), v(*), w(*), p(*)
    end function
    integer(c_int) function pf_download(handle, u, v, w, p) bind(C, name="pf_download")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(inout) :: u(*), v(*), w(*), p(*)
    end function
    integer(c_int) function pf_step(handle, nsteps, p_error) bind(C, name="pf_step")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: nsteps
      real(c_double), intent(out) :: p_error(*)
    end function
    integer(c_int) function pf_step_host(handle, nsteps, u, v, w, p, p_error) bind(C, name="pf_step_host")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: nsteps
      real(c_double), intent(inout) :: u(*), v(*), w(*), p(*)
      real(c_double), intent(out) :: p_error(*)
    end function
    integer(c_int) function pf_initial_conditions(handle) bind(C, name="pf_initial_conditions")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_copy_old(handle) bind(C, name="pf_copy_old")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_divergence(handle) bind(C, name="pf_divergence")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_predictor(handle) bind(C, name="pf_predictor")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_build_poisson(handle) bind(C, name="pf_build_poisson")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_sor(handle, iters, p_error) bind(C, name="pf_sor")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: iters
      real(c_double), intent(out) :: p_error
    end function
    integer(c_int) function pf_project(handle) bind(C, name="pf_project")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_boundary(handle) bind(C, name="pf_boundary")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    ! output_force_log_2d (lib/output.f90:244-305): out8 = Fp_x, Fp_y, Fv_x, Fv_y, F_x, F_y, Cd, Cl
    integer(c_int) function pf_force_log_2d(handle, radius, out8) bind(C, name="pf_force_log_2d")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: radius
      real(c_double), intent(out) :: out8(8)
    end function
    ! output_force_log_3d (lib/output.f90:1090-1165): out12 = Fp xyz, Fv xyz, F xyz, Cd(x), Cl, Cd(z)
    integer(c_int) function pf_force_log_3d(handle, radius, out12) bind(C, name="pf_force_log_3d")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: radius
      real(c_double), intent(out) :: out12(12)
    end function
    ! bodies of the VTK snapshot sections (lib/output.f90:968-1088 / :421-537), formatted "(3(f16.4,1x))" on the GPU:
    !   allocate(character(len=pf_vtk_section_bytes(h, sec, nk)) :: text); ierr = pf_vtk_section(h, sec, k0, nk, xp, yp, zp, text)
    !   write(65) text   ! unit opened with access="stream"
    integer(c_size_t) function pf_vtk_section_bytes(handle, section, nplanes) bind(C, name="pf_vtk_section_bytes")
      import :: c_size_t, c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: section, nplanes
    end function
    integer(c_int) function pf_vtk_section(handle, section, k_local0, nplanes, xp, yp, zp, text) &
        bind(C, name="pf_vtk_section")
      import :: c_int, c_ptr, c_double, c_char
      type(c_ptr), value :: handle
      integer(c_int), value :: section, k_local0, nplanes
      real(c_double), intent(in) :: xp(*), yp(*), zp(*)
      character(kind=c_char), intent(out) :: text(*)
    end function
    ! z-slab ranks forked from this program (pf_ranks.cu).  nranks = 0: take the count from the environment variable
    ! PIXELFLOW_GPUS (default 1) -- the analogue of OMP_NUM_THREADS in the reference's config/omp_config.conf
    integer(c_int) function pf_ranks_launch(nranks, rank) bind(C, name="pf_ranks_launch")
      import :: c_int
      integer(c_int), value :: nranks
      integer(c_int), intent(out) :: rank
    end function
    integer(c_int) function pf_ranks_rank() bind(C, name="pf_ranks_rank")
      import :: c_int
    end function
    integer(c_int) function pf_ranks_count() bind(C, name="pf_ranks_count")
      import :: c_int
    end function
    type(c_ptr) function pf_ranks_unique_id() bind(C, name="pf_ranks_unique_id")
      import :: c_ptr
    end function
    integer(c_int) function pf_ranks_barrier() bind(C, name="pf_ranks_barrier")
      import :: c_int
    end function
    integer(c_int) function pf_ranks_finish(status) bind(C, name="pf_ranks_finish")
      import :: c_int
      integer(c_int), value :: status
    end function
    ! the whole fields in rank 0's arrays (collective; on one rank = pf_download)
    integer(c_int) function pf_gather(handle, u, v, w, p) bind(C, name="pf_gather")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: u(*), v(*), w(*), p(*)
    end function
    integer(c_int) function pf_sync(handle) bind(C, name="pf_sync")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function pf_last_timing(handle, ms_total, ms_sor, launches) bind(C, name="pf_last_timing")
      import :: c_int, c_ptr, c_double, c_long_long
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: ms_total, ms_sor
      integer(c_long_long), intent(out) :: launches
    end function
  end interface

contains

  ! C string returned by pf_last_error -> Fortran string
  function pf_error_message(handle) result(msg)
    type(c_ptr), intent(in) :: handle
    character(len=:), allocatable :: msg
    type(c_ptr) :: cp
    character(kind=c_char), pointer :: chars(:)
    integer :: n
    cp = pf_last_error(handle)
    msg = ''
    if (.not. c_associated(cp)) return
    call c_f_pointer(cp, chars, [1024])
    n = 0
    do while (n < 1024)
      if (chars(n + 1) == c_null_char) exit
      n = n + 1
    end do
    allocate (character(len=n) :: msg)
    msg = transfer(chars(1:n), msg)
  end function pf_error_message

  ! the error convention of the ABI: every call returns 0 on success; stop like the reference would crash
  subroutine pf_check(ierr, handle, what)
    integer(c_int), intent(in) :: ierr
    type(c_ptr), intent(in) :: handle
    character(len=*), intent(in) :: what
    if (ierr /= 0) then
      write (*, *) 'pixelflow_gpu: ', what, ' failed: ', pf_error_message(handle)
      stop 1
    end if
  end subroutine pf_check

end module pixelflow_gpu
