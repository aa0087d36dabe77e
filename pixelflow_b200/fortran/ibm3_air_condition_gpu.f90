! ibm3_air_condition_gpu.f90 -- a PixelFlow ibm3_air_condition driver (3D room, six wall / inlet / outlet faces) whose
! time loop runs on the GPU.
!
! Written against the reference's support library (global_3d, valiables, grid_3d, output_3d, utils:
! src/omp_parallel/lib/*.f90), which it links unchanged apart from raising md,nd,ld in lib/global.f90
! to the grid at hand.  Inputs (config/controlDict.txt namelists, porosity CSV), log lines and output
! files are the reference's; everything between "MAC algorithm start" and the final outputs is
! libpixelflow_gpu.so.  Build (not possible in this repository's image, which has no Fortran compiler):
!
!   gfortran -O2 -fdefault-real-8 -fno-automatic -mcmodel=medium \
!       lib/global.f90 lib/utils.f90 lib/grid.f90 lib/output.f90 \
!       pixelflow_gpu_mod.f90 ibm3_air_condition_gpu.f90 -L<repo>/pixelflow_b200 -lpixelflow_gpu -o ibm3_air_condition_omp
program ibm3_air_condition_gpu
  use iso_c_binding
  use global_3d
  use valiables
  use output_3d
  use grid_3d
  use utils
  use pixelflow_gpu
  implicit none
  real, dimension(0:md, 0:nd, 0:ld) :: u, v, w, p, porosity
  real, dimension(0:md) :: xp
  real, dimension(0:nd) :: yp
  real, dimension(0:ld) :: zp
  real :: dx, dy, dz, dt, p_error(1)
  integer :: m, n, l, istep
  integer(c_int) :: rank, nranks, istat
  integer, parameter :: top_wall = 1, bottom_wall = 0, east_wall = 0, west_wall = 0, south_wall = 2, north_wall = 0
  type(pf_config) :: cfg
  type(c_ptr) :: h

  call get_now_time()
  m = 0
  density = 0.
  call read_settings(xnue, xlambda, density, width, height, depth, time, inlet_velocity, outlet_pressure, AoA, &
                     istep_max, istep_out, thickness, threshold, radius, center_x, center_y, center_z, &
                     nonslip, output_folder, csv_file, iter_max, relux_factor)
  call system('mkdir -p '//trim(output_folder))
  call system('mkdir -p etc')
  call grid_conditions_wall(xp, yp, zp, dx, dy, dz, dt, xnue, xlambda, density, width, height, depth, &
                          thickness, threshold, radius, center_x, center_y, center_z, time, &
                          inlet_velocity, AoA, porosity, m, n, l, istep_max, csv_file)
  call output_grid_3d(xp, yp, zp, m, n, l)
  write (*, *) '# istep_max= ', istep_max, '   istep_out= ', istep_out

  ! ---- several GPUs: PIXELFLOW_GPUS copies of this program from here on (one per GPU, z-slabs of the grid) --------
  ! The deck is read, the grid file written; nothing has touched CUDA yet.  Rank 0 is this process and keeps the
  ! log; the other ranks inherit the arrays read above and write nothing (their stdout is discarded, and every file
  ! below is written under `if (rank == 0)`).  With PIXELFLOW_GPUS unset or 1 nothing happens here.
  call pf_check(pf_ranks_launch(0, rank), c_null_ptr, 'pf_ranks_launch')
  nranks = pf_ranks_count()

  ! ---- hand the problem to the GPU library -------------------------------------------------------
  call pf_config_init(cfg)
  cfg%solver_case = PF_IBM3_AIRCOND
  ! module wall_conditions of the reference (ibm_3d_air_condition_omp_cpu.f90:4-16): compile-time there, run-time
  ! here.  0 wall, 1 inlet where porosity >= 0.9, 2 outlet where porosity >= 0.9
  cfg%wall(PF_TOP) = top_wall
  cfg%wall(PF_BOTTOM) = bottom_wall
  cfg%wall(PF_EAST) = east_wall
  cfg%wall(PF_WEST) = west_wall
  cfg%wall(PF_SOUTH) = south_wall
  cfg%wall(PF_NORTH) = north_wall
  cfg%m = m; cfg%n = n; cfg%l = l
  cfg%host_ldx = md + 1            ! the arrays are dimension(0:md,0:nd,0:ld)
  cfg%host_ldy = nd + 1
  cfg%dx = dx; cfg%dy = dy; cfg%dz = dz; cfg%dt = dt
  cfg%xnue = xnue; cfg%xlambda = xlambda; cfg%density = density; cfg%thickness = thickness
  cfg%nonslip = merge(1, 0, nonslip)
  cfg%iter_max = iter_max
  cfg%relux_factor = relux_factor
  cfg%inlet_velocity = inlet_velocity; cfg%outlet_pressure = outlet_pressure; cfg%AoA = AoA
  cfg%rank = rank
  cfg%nranks = nranks
  if (nranks > 1) then
    cfg%device = rank                         ! one GPU per rank
    cfg%nccl_unique_id = pf_ranks_unique_id() ! made by rank 0, awaited by the others
  end if
  if (pf_create(h, cfg) /= 0) then
    write (*, *) 'pixelflow_gpu: pf_create failed: ', pf_error_message(c_null_ptr)
    stop 1
  end if
  call pf_check(pf_set_porosity(h, porosity), h, 'pf_set_porosity')
  u = 0.; v = 0.; w = 0.; p = 0.
  call pf_check(pf_upload(h, u, v, w, p), h, 'pf_upload')
  call pf_check(pf_initial_conditions(h), h, 'pf_initial_conditions')   ! initial_conditions + boundary
  call pf_check(pf_gather(h, u, v, w, p), h, 'pf_gather')          ! the whole fields in rank 0's arrays
  if (rank == 0) call output_paraview_temp_3d(p, u, v, w, porosity, xp, yp, zp, m, n, l, 0)

  call get_now_time()
  write (*, *) '# --- MAC algorithm start'
  do istep = 1, istep_max
    time = istep*dt
    write (*, *) '--- time_steps= ', istep, ' --  time = ', time
    call pf_check(pf_step(h, 1, p_error), h, 'pf_step')   ! u_old copy, solve_p, projection, boundary
    write (*, *) 'SOR iteration no.', iter_max, '-- p error:', p_error(1)
    if (mod(istep, istep_out) == 0) then
      call pf_check(pf_gather(h, u, v, w, p), h, 'pf_gather')
      if (rank == 0) call output_paraview_temp_3d(p, u, v, w, porosity, xp, yp, zp, m, n, l, istep)
    end if
  end do
  call get_now_time()

  call pf_check(pf_gather(h, u, v, w, p), h, 'pf_gather')
  call pf_destroy(h)
  istat = pf_ranks_finish(0)                   ! ranks > 0 end here; rank 0 goes on once they have
  if (istat /= 0) then
    write (*, *) 'pixelflow_gpu: a GPU rank failed'
    stop 1
  end if
  call output_solution_post_3d(p, u, v, w, xp, yp, zp, porosity, m, n, l)
  call output_divergent_3d(p, u, v, w, porosity, dx, dy, dz, m, n, l)
  call output_paraview_3d(p, u, v, w, porosity, xp, yp, zp, m, n, l)
  write (*, *) 'program finished'
  call get_now_time()
end program ibm3_air_condition_gpu
