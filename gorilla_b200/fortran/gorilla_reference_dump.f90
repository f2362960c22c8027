!> gorilla_reference_dump -- reference dump shim (SURVEY.md section 7 step 9, section 8c pin (5)).
!!
!! A small driver that links against the UNMODIFIED GORILLA library (libGORILLA.a of the reference's own build)
!! and writes, for the settings of the gorilla.inp / tetra_grid.inp in the working directory,
!!   (1) the mesh the reference built:  tetra_physics(1:ntetr), tetra_grid(1:ntetr), the vertex tables and the
!!       module scalars the hot path reads (what `gorilla_mesh_desc` of include/gorilla_b200.h carries), and
!!   (2) for the particles of `dump_particles.bin`: the state after n_steps successive orbit_timestep_gorilla calls (the
!!       first one locates the particle, the later ones start from ind_tetr / iface) and the (ind_tetr, iface) pair after
!!       each of the first trace_cap pusher calls,
!! into one little-endian stream file `gorilla_reference_dump.bin` (layout below = tests/reference_dump.py, which
!! reads it back, turns it into a .gmesh + golden vectors and checks the C oracle and the CUDA path against it).
!! This is what turns "parity unpinned" into "pinned against the gfortran binary".
!!
!! The time-step loop below is the loop of orbit_timestep_gorilla (SRC/orbit_timestep_gorilla.f90:46-144) written
!! against the reference's PUBLIC procedures (check_coordinate_domain, find_tetra, bmod_func, vperp_func,
!! initialize_const_motion_*, pusher_tetra_poly / pusher_tetra_rk); the only addition is the recording of
!! (ind_tetr, iface) after each pusher call.  The final state is cross-checked in the program itself against a
!! plain call of orbit_timestep_gorilla on a copy of the same particle (n_mismatch in the log must be 0).
!!
!! Build (on a machine with gfortran + NetCDF-Fortran + LAPACK, after the reference's own cmake build):
!!     gfortran -O2 -I<GORILLA>/BUILD/OBJS gorilla_reference_dump.f90 -L<GORILLA>/BUILD -lGORILLA \
!!              -lnetcdff -lnetcdf -llapack -fopenmp -o gorilla_reference_dump.x
!! Run (in a directory holding gorilla.inp, tetra_grid.inp, the equilibrium links of an EXAMPLES/ folder and a
!! dump_particles.bin made by `python tests/reference_dump.py particles ...`):
!!     ./gorilla_reference_dump.x
!!
!! NOTE: this image has no Fortran compiler (SURVEY.md F2); the program is shipped as source, its file format is
!! exercised by tests/test_reference_dump.py through a Python writer of the same layout.
!!
!! File layout (all little endian, no record markers):
!!   char[8]  'GREFDMP1'
!!   int32    ndoubles_per_tetra (142), nints_per_tetra (20), ntetr, nvert, has_sthetaphi, has_skew,
!!            sign_sqg, coord_system, n_field_periods, grid_kind, grid_size(3)                      [13]
!!   int32    ispecies, boole_periodic_relocation, ipusher, boole_pusher_ode45, boole_dt_dtau, boole_newton_precalc,
!!            poly_order, i_precomp, boole_guess, i_time_tracing_option, handover_processing_kind,
!!            boole_adaptive_time_steps, boole_strong_electric_field, max_n_intermediate_steps      [14]
!!   real64   cm_over_e, particle_mass, particle_charge, Rmin, Rmax, Zmin, Zmax, sfc_s_min, eps_Phi,
!!            desired_delta_energy                                                                  [10]
!!   real64   tetra_physics  [ntetr][142]        int32 tetra_grid [ntetr][20]
!!   real64   verts_rphiz [nvert][3]             real64 verts_sthetaphi [nvert][3]   (if has_sthetaphi)
!!   real64   tetra_skew_coord [ntetr][168]      (if has_skew)
!!   int32    n_particles, trace_cap, n_steps    real64 t_step
!!   real64   x0 [n][3], vpar0 [n], vperp0 [n]                     (the inputs, echoed)
!!   real64   x [n][3], vpar [n], vperp [n], t_remain [n]          (after the last call made for the particle; a particle
!!                                                                  that was not placed or has left the domain is not
!!                                                                  passed to orbit_timestep_gorilla again)
!!   int32    boole_initialized [n], ind_tetr [n], iface [n], n_pushes [n]   (n_pushes: sum over the calls)
!!   int32    trace_ind_tetr [n][trace_cap], trace_iface [n][trace_cap]        (unused slots 0)
program gorilla_reference_dump
  use tetra_grid_settings_mod, only: load_tetra_grid_inp, grid_kind, grid_size, n_field_periods, sfc_s_min
  use gorilla_settings_mod, only: load_gorilla_inp, eps_Phi, ispecies, boole_periodic_relocation, ipusher, boole_pusher_ode45, &
                                  boole_dt_dtau, boole_newton_precalc, poly_order, i_precomp, boole_guess, &
                                  i_time_tracing_option, handover_processing_kind, boole_adaptive_time_steps, &
                                  desired_delta_energy, max_n_intermediate_steps, boole_strong_electric_field
  use orbit_timestep_gorilla_mod, only: initialize_gorilla, orbit_timestep_gorilla, check_coordinate_domain
  use tetra_physics_mod, only: tetra_physics, tetra_skew_coord, cm_over_e, particle_mass, particle_charge, sign_sqg, &
                               cs_phys => coord_system
  use tetra_grid_mod, only: tetra_grid, ntetr, nvert, verts_rphiz, verts_sthetaphi, Rmin, Rmax, Zmin, Zmax
  use find_tetra_mod, only: find_tetra
  use supporting_functions_mod, only: bmod_func, vperp_func
  use pusher_tetra_poly_mod, only: pusher_tetra_poly, initialize_const_motion_poly
  use pusher_tetra_rk_mod, only: pusher_tetra_rk, initialize_const_motion_rk
  use, intrinsic :: iso_fortran_env, only: int32, real64
  implicit none

  integer :: u, n, cap, i, k, n_steps, n_mismatch
  integer(int32) :: n32, cap32, nsteps32, np1, has_sthetaphi, has_skew
  real(real64) :: t_step
  real(real64), allocatable :: x0(:,:), vpar0(:), vperp0(:), x(:,:), vpar(:), vperp(:), t_rem(:)
  integer(int32), allocatable :: binit(:), itetr(:), ifc(:), npush(:), tr_tetr(:,:), tr_face(:,:)
  ! scalar copies for the cross-check against the unmodified entry point
  real(real64) :: xc(3), vparc, vperpc
  logical :: binitc
  integer :: itetrc, ifcc

  call load_tetra_grid_inp()
  call load_gorilla_inp()
  call initialize_gorilla()

  ! ---- particles -------------------------------------------------------------------------------------------
  open(newunit=u, file='dump_particles.bin', access='stream', form='unformatted', status='old', action='read')
  read(u) n32, cap32, nsteps32, t_step
  n = n32; cap = max(int(cap32), 1); n_steps = max(int(nsteps32), 1)
  allocate(x0(3,n), vpar0(n), vperp0(n), x(3,n), vpar(n), vperp(n), t_rem(n))
  allocate(binit(n), itetr(n), ifc(n), npush(n), tr_tetr(cap,n), tr_face(cap,n))
  read(u) x0, vpar0, vperp0
  close(u)
  x = x0; vpar = vpar0; vperp = vperp0
  t_rem = 0.d0; binit = 0; itetr = -1; ifc = -1; npush = 0; tr_tetr = 0; tr_face = 0

  ! The pushers keep particle-private state in threadprivate module variables; the loop stays serial so that
  ! the dump does not depend on the OpenMP schedule.
  n_mismatch = 0
  do i = 1, n
    do k = 1, n_steps
      call traced_timestep(x(:,i), vpar(i), vperp(i), t_step, binit(i), itetr(i), ifc(i), t_rem(i), np1, npush(i), &
                           tr_tetr(:,i), tr_face(:,i))
      npush(i) = npush(i) + np1
      if (binit(i) == 0 .or. itetr(i) == -1) exit     ! not placed / left the domain: no further calls
    end do
    xc = x0(:,i); vparc = vpar0(i); vperpc = vperp0(i); binitc = .false.; itetrc = -1; ifcc = -1
    do k = 1, n_steps
      call orbit_timestep_gorilla(xc, vparc, vperpc, t_step, binitc, itetrc, ifcc)
      if (.not. binitc .or. itetrc == -1) exit
    end do
    if (any(xc /= x(:,i)) .or. vparc /= vpar(i) .or. vperpc /= vperp(i) .or. itetrc /= itetr(i) .or. ifcc /= ifc(i)) &
      n_mismatch = n_mismatch + 1
  end do
  print *, 'gorilla_reference_dump: particles = ', n, ' n_mismatch (traced loop vs orbit_timestep_gorilla) = ', n_mismatch

  ! ---- file ------------------------------------------------------------------------------------------------
  has_sthetaphi = 0     ! only the field-aligned grids allocate it (create_points)
  if (allocated(verts_sthetaphi)) then
    if (size(verts_sthetaphi, 2) >= nvert) has_sthetaphi = 1
  end if
  has_skew = merge(1, 0, handover_processing_kind == 2)
  open(newunit=u, file='gorilla_reference_dump.bin', access='stream', form='unformatted', status='replace', action='write')
  write(u) 'GREFDMP1'
  write(u) 142_int32, 20_int32, int(ntetr, int32), int(nvert, int32), has_sthetaphi, has_skew, &
           int(sign_sqg, int32), int(cs_phys, int32), int(n_field_periods, int32), int(grid_kind, int32), &
           int(grid_size, int32)
  write(u) int(ispecies, int32), l2i(boole_periodic_relocation), int(ipusher, int32), l2i(boole_pusher_ode45), &
           l2i(boole_dt_dtau), l2i(boole_newton_precalc), int(poly_order, int32), int(i_precomp, int32), &
           l2i(boole_guess), int(i_time_tracing_option, int32), int(handover_processing_kind, int32), &
           l2i(boole_adaptive_time_steps), l2i(boole_strong_electric_field), int(max_n_intermediate_steps, int32)
  write(u) cm_over_e, particle_mass, particle_charge, Rmin, Rmax, Zmin, Zmax, sfc_s_min, eps_Phi, desired_delta_energy
  write(u) tetra_physics(1:ntetr)      ! `sequence` type of 142 doubles: the bytes as they sit in memory
  write(u) tetra_grid(1:ntetr)         ! `sequence` type of 20 default integers
  write(u) verts_rphiz(:, 1:nvert)
  if (has_sthetaphi == 1) write(u) verts_sthetaphi(:, 1:nvert)
  if (has_skew == 1) write(u) tetra_skew_coord(1:ntetr)
  write(u) n32, int(cap, int32), int(n_steps, int32), t_step
  write(u) x0, vpar0, vperp0
  write(u) x, vpar, vperp, t_rem
  write(u) binit, itetr, ifc, npush
  write(u) tr_tetr, tr_face
  close(u)
  print *, 'gorilla_reference_dump: wrote gorilla_reference_dump.bin, ntetr = ', ntetr
  if (n_mismatch /= 0) stop 1

contains

  pure integer(int32) function l2i(b)
    logical, intent(in) :: b
    l2i = merge(1_int32, 0_int32, b)
  end function

  !> One orbit_timestep_gorilla call, recording the cell/face after every push: push number n_before + j of the particle
  !> goes to slot n_before + j of the trace (n_before = pushes of its earlier calls).
  subroutine traced_timestep(xp, vparp, vperpp, dt, binit_p, ind, face, t_remain, n_push, n_before, trace_t, trace_f)
    real(real64), intent(inout) :: xp(3), vparp, vperpp
    real(real64), intent(in) :: dt
    integer(int32), intent(inout) :: binit_p, ind, face
    real(real64), intent(out) :: t_remain
    integer(int32), intent(out) :: n_push
    integer(int32), intent(in) :: n_before
    integer(int32), intent(inout) :: trace_t(:), trace_f(:)
    real(real64) :: z_save(3), perpinv, perpinv2, t_pass
    logical :: finished
    integer :: ind_l, face_l, ind_save, iper

    n_push = 0; t_remain = 0.d0
    ind_l = ind; face_l = face
    if (binit_p == 0) then
      call check_coordinate_domain(xp)
      call find_tetra(xp, vparp, vperpp, ind_l, face_l, int(sign(1.d0, dt)))
      ind = ind_l; face = face_l
      if (ind_l == -1) return
      binit_p = 1
    end if
    if (dt == 0.d0) return

    z_save = xp - tetra_physics(ind_l)%x1
    perpinv = -0.5d0*vperpp**2/bmod_func(z_save, ind_l)
    perpinv2 = perpinv**2
    if (ipusher == 1) then
      call initialize_const_motion_rk(perpinv, perpinv2)
    else
      call initialize_const_motion_poly(perpinv, perpinv2)
    end if

    t_remain = dt
    finished = .false.
    ind_save = ind_l
    do while (ind_l /= -1)
      ind_save = ind_l
      if (ipusher == 1) then
        call pusher_tetra_rk(ind_l, face_l, xp, vparp, z_save, t_remain, t_pass, finished, iper)
      else
        call pusher_tetra_poly(poly_order, ind_l, face_l, xp, vparp, z_save, t_remain, t_pass, finished, iper)
      end if
      n_push = n_push + 1
      if (n_before + n_push <= size(trace_t)) then
        trace_t(n_before + n_push) = ind_l; trace_f(n_before + n_push) = face_l
      end if
      t_remain = t_remain - t_pass
      if (finished) exit
    end do
    vperpp = vperp_func(z_save, perpinv, ind_save)
    ind = ind_l; face = face_l
  end subroutine

end program gorilla_reference_dump
