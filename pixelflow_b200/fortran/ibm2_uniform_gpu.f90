! ibm2_uniform_gpu.f90 -- a PixelFlow ibm2_uniform_omp driver whose time loop runs on the GPU.
!
! Same construction as ibm3_uniform_gpu.f90: the reference's support library (global_2d, valiables, grid_2d,
! output_2d, utils: src/omp_parallel/lib/*.f90) is linked unchanged apart from md, nd in lib/global.f90; inputs, log
! lines and output files are the reference's; everything between "MAC algorithm start" and the final outputs is
! libpixelflow_gpu.so (src/omp_parallel/ibm_2d_uniform_omp_cpu.f90:80-125).
!
!   gfortran -O2 -fdefault-real-8 -fno-automatic -mcmodel=medium \
!       lib/global.f90 lib/utils.f90 lib/grid.f90 lib/output.f90 \
!       pixelflow_gpu_mod.f90 ibm2_uniform_gpu.f90 -L<repo>/pixelflow_b200 -lpixelflow_gpu -o ibm2_uniform_omp
program ibm2_uniform_gpu
  use iso_c_binding
  use global_2d
  use valiables
  use output_2d
  use grid_2d
  use utils
  use pixelflow_gpu
  implicit none
  real, dimension(0:md, 0:nd) :: u, v, p, porosity
  real, dimension(1) :: w_unused      ! the 2D cases have no w: pf_upload / pf_download ignore the argument
  real, dimension(0:md) :: xp
  real, dimension(0:nd) :: yp
  real :: dx, dy, dt, p_error(1)
  integer :: m, n, istep
  type(pf_config) :: cfg
  type(c_ptr) :: h

  call get_now_time()
  call read_settings(xnue, xlambda, density, width, height, depth, time, inlet_velocity, outlet_pressure, AoA, &
                     istep_max, istep_out, thickness, threshold, radius, center_x, center_y, center_z, &
                     nonslip, output_folder, csv_file, iter_max, relux_factor)
  call system('mkdir -p '//trim(output_folder))
  call system('mkdir -p etc')
  call grid_conditions(xp, yp, dx, dy, dt, xnue, xlambda, density, width, height, depth, &
                       thickness, threshold, radius, center_x, center_y, time, &
                       inlet_velocity, AoA, porosity, m, n, istep_max, csv_file)
  call output_grid_2d(xp, yp, m, n)
  write (*, *) '# istep_max= ', istep_max, '   istep_out= ', istep_out

  ! ---- hand the problem to the GPU library -------------------------------------------------------
  call pf_config_init(cfg)
  cfg%solver_case = PF_IBM2_UNIFORM
  cfg%m = m; cfg%n = n; cfg%l = 1
  cfg%host_ldx = md + 1            ! the arrays are dimension(0:md,0:nd)
  cfg%host_ldy = nd + 1
  cfg%dx = dx; cfg%dy = dy; cfg%dt = dt
  cfg%xnue = xnue; cfg%xlambda = xlambda; cfg%density = density; cfg%thickness = thickness
  cfg%nonslip = merge(1, 0, nonslip)
  cfg%iter_max = iter_max
  cfg%relux_factor = relux_factor
  cfg%inlet_velocity = inlet_velocity; cfg%outlet_pressure = outlet_pressure; cfg%AoA = AoA
  if (pf_create(h, cfg) /= 0) then
    write (*, *) 'pixelflow_gpu: pf_create failed: ', pf_error_message(c_null_ptr)
    stop 1
  end if
  call pf_check(pf_set_porosity(h, porosity), h, 'pf_set_porosity')
  u = 0.; v = 0.; p = 0.
  call pf_check(pf_upload(h, u, v, w_unused, p), h, 'pf_upload')
  call pf_check(pf_initial_conditions(h), h, 'pf_initial_conditions')   ! initial_conditions + boundary
  call pf_check(pf_download(h, u, v, w_unused, p), h, 'pf_download')
  call output_paraview_temp_2d(p, u, v, porosity, xp, yp, m, n, inlet_velocity, 0, output_folder)

  call get_now_time()
  write (*, *) '# --- MAC algorithm start'
  do istep = 1, istep_max
    time = istep*dt
    write (*, *) '--- time_steps= ', istep, ' --  time = ', time
    call pf_check(pf_step(h, 1, p_error), h, 'pf_step')   ! u_old copy, solve_p, projection, boundary
    write (*, *) 'SOR iteration no.', iter_max, '-- p error:', p_error(1)
    if (mod(istep, istep_out) == 0) then
      call pf_check(pf_download(h, u, v, w_unused, p), h, 'pf_download')
      call output_paraview_temp_2d(p, u, v, porosity, xp, yp, m, n, inlet_velocity, istep, output_folder)
    end if
  end do
  call get_now_time()

  call pf_check(pf_download(h, u, v, w_unused, p), h, 'pf_download')
  call pf_destroy(h)
  call output_solution_post_2d(p, u, v, xp, yp, porosity, m, n)
  call output_divergent_2d(p, u, v, porosity, dx, dy, m, n)
  call output_paraview_2d(p, u, v, porosity, xp, yp, m, n, inlet_velocity, output_folder)
  write (*, *) 'program finished'
  call get_now_time()
end program ibm2_uniform_gpu
