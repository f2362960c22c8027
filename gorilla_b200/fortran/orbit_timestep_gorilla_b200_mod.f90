!> orbit_timestep_gorilla_b200_mod -- Fortran side of the drop-in boundary (ISO_C_BINDING).
!!
!! Batched driver for GORILLA's particle-parallel hot path.  It keeps `orbit_timestep_gorilla` as the entry
!! point (same argument list as SRC/orbit_timestep_gorilla.f90:19) and adds `orbit_timestep_gorilla_batch`;
!! both forward to the C ABI of libgorilla_b200.so (include/gorilla_b200.h), whose kernels run on the GPU.
!!
!! Usage inside GORILLA (after the usual host initialisation, which stays unchanged):
!!     call load_tetra_grid_inp(); call load_gorilla_inp(); call initialize_gorilla()
!!     call initialize_gorilla_b200()                       ! uploads tetra_physics / tetra_grid once
!!     call orbit_timestep_gorilla_batch(n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, ierr)
!!
!! Multi-GPU (one process per GPU, particles sharded, mesh replicated on every GPU):
!!     if (rank == 0) call comm_unique_id_b200(id)          ! 128 bytes
!!     call MPI_Bcast(id, 128, MPI_BYTE, 0, comm, ierr)     ! or any other way to hand them to the ranks
!!     call comm_init_b200(id, rank, nranks, ierr)
!!     call shard_range_b200(n_total, rank, nranks, first, count)      ! this rank pushes particles first+1 .. first+count
!!     ... time steps ...
!!     call diag_reduce_b200(n, x, vpar, vperp, ind_tetr, e0, pphi0, perpinv0, diag, ierr)   ! counters + conservation, all ranks
!!
!! NOTE: this image has no Fortran compiler (SURVEY.md F2), so this module is shipped as source and is not
!! part of the automated build; the C ABI it binds is exercised by the Python ctypes binding in the tests.
module orbit_timestep_gorilla_b200_mod
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: initialize_gorilla_b200, finalize_gorilla_b200, orbit_timestep_gorilla, orbit_timestep_gorilla_batch, &
            orbit_timestep_gorilla_batch_optional, orbit_timestep_gorilla_batch_events, find_tetra_batch, &
            gorilla_b200_counters_t, get_counters_b200, invariants_b200, set_host_resort_b200, set_gather_b200, &
            gorilla_b200_diag_t, gorilla_b200_event_t, gorilla_b200_event_settings_t, &
            comm_unique_id_b200, comm_init_b200, comm_free_b200, shard_range_b200, diag_reset_b200, diag_reduce_b200

  !> struct gorilla_settings (include/gorilla_b200.h)
  type, bind(C) :: gorilla_settings_t
    real(c_double)  :: eps_Phi
    integer(c_int32_t) :: coord_system, ispecies, boole_periodic_relocation, ipusher, boole_pusher_ode45, &
                          boole_dt_dtau, boole_newton_precalc, poly_order, i_precomp, boole_guess, &
                          i_time_tracing_option, handover_processing_kind, boole_adaptive_time_steps, &
                          boole_strong_electric_field, boole_grid_for_find_tetra, &
                          boole_time_Hamiltonian, boole_gyrophase, boole_vpar_int, boole_vpar2_int, &
                          max_n_intermediate_steps
    real(c_double)  :: desired_delta_energy, rel_err_ode45
    ! read by gorilla_mesh_build only (a Fortran caller's make_tetra_physics has already applied the perturbation)
    real(c_double)  :: helical_pert_eps_Aphi = 0.d0
    integer(c_int32_t) :: boole_helical_pert = 0, helical_pert_m_fourier = 0, helical_pert_n_fourier = 0, reserved0 = 0
    real(c_double)  :: axi_noise_eps_A = 0.d0, axi_noise_eps_Phi = 0.d0, non_axi_noise_eps_A = 0.d0
    integer(c_int32_t) :: boole_axi_noise_vector_pot = 0, boole_axi_noise_elec_pot = 0, boole_non_axi_noise_vector_pot = 0, &
                          noise_seed = 0
  end type
  !> struct gorilla_mesh_desc
  type, bind(C) :: gorilla_mesh_desc_t
    integer(c_int64_t) :: ntetr
    type(c_ptr)        :: tetra_physics, tetra_grid
    real(c_double)     :: cm_over_e, particle_mass, particle_charge
    integer(c_int32_t) :: sign_sqg, coord_system, n_field_periods, grid_kind
    integer(c_int32_t) :: grid_size(3), pad0
    real(c_double)     :: Rmin, Rmax, Zmin, Zmax, sfc_s_min
    type(c_ptr)        :: tetra_skew_coord   ! c_loc(tetra_skew_coord(1)) if handover_processing_kind = 2, else c_null_ptr
  end type
  !> struct gorilla_counters
  type, bind(C) :: gorilla_b200_counters_t
    integer(c_int64_t) :: n_particles, n_pushes, n_lost, n_finished, n_fallback(4), n_domain_errors
    real(c_double)     :: kernel_ms, find_ms
    integer(c_int64_t) :: n_adaptive, n_lost_inner, n_failed
  end type
  !> struct gorilla_diag: counters since diag_reset_b200 and conservation statistics, reduced over all ranks
  type, bind(C) :: gorilla_b200_diag_t
    integer(c_int64_t) :: n_particles, n_pushes, n_lost, n_lost_outer, n_lost_inner, n_failed, n_finished
    integer(c_int64_t) :: n_fallback(4), n_adaptive, n_sampled
    real(c_double)     :: max_delta_energy, rms_delta_energy, max_delta_perpinv, rms_delta_perpinv
    real(c_double)     :: max_delta_p_phi, rms_delta_p_phi
    integer(c_int32_t) :: nranks, reserved
  end type
  !> struct gorilla_event / gorilla_event_settings (banana tips with J_par, toroidal mappings; gorilla_plot_mod.f90:585-638)
  type, bind(C) :: gorilla_b200_event_t
    integer(c_int64_t) :: particle
    integer(c_int32_t) :: kind, counter
    integer(c_int64_t) :: push
    real(c_double)     :: x(3), value(2)
    real(c_double)     :: t   ! t_step - t_remain after the push of the event
  end type
  type, bind(C) :: gorilla_b200_event_settings_t
    integer(c_int32_t) :: boole_poincare_phi_0, n_skip_phi_0, boole_poincare_vpar_0, boole_J_par, n_skip_vpar_0
    integer(c_int32_t) :: boole_full_orbit = 0, n_skip_full_orbit = 1   ! kind 3 events: the full_orbit_plot / p_phi / e_tot files
    integer(c_int32_t) :: reserved = 0
  end type

  interface
    integer(c_int) function gorilla_b200_init(mesh, settings, handle) bind(C, name='gorilla_b200_init')
      import :: c_int, c_ptr, gorilla_mesh_desc_t, gorilla_settings_t
      type(gorilla_mesh_desc_t), intent(in) :: mesh
      type(gorilla_settings_t), intent(in)  :: settings
      type(c_ptr), intent(out)              :: handle
    end function
    subroutine gorilla_b200_free(handle) bind(C, name='gorilla_b200_free')
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine
    integer(c_int) function gorilla_b200_orbit_timestep(handle, n, x, vpar, vperp, t_step, boole_initialized, &
                                                        ind_tetr, iface, t_remain_out, n_pushes) &
                                                        bind(C, name='gorilla_b200_orbit_timestep')
      import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t
      type(c_ptr), value        :: handle
      integer(c_int64_t), value :: n
      real(c_double)            :: x(3,*), vpar(*), vperp(*)
      real(c_double), value     :: t_step
      integer(c_int32_t)        :: boole_initialized(*), ind_tetr(*), iface(*)
      type(c_ptr), value        :: t_remain_out, n_pushes      ! c_null_ptr if not wanted
    end function
    integer(c_int) function gorilla_b200_find_tetra(handle, n, x, vpar, vperp, ind_tetr, iface, sign_t_step) &
                                                    bind(C, name='gorilla_b200_find_tetra')
      import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t
      type(c_ptr), value        :: handle
      integer(c_int64_t), value :: n
      real(c_double)            :: x(3,*)
      real(c_double), intent(in):: vpar(*), vperp(*)
      integer(c_int32_t)        :: ind_tetr(*), iface(*)
      integer(c_int32_t), value :: sign_t_step
    end function
    integer(c_int) function gorilla_b200_get_counters(handle, counters) bind(C, name='gorilla_b200_get_counters')
      import :: c_int, c_ptr, gorilla_b200_counters_t
      type(c_ptr), value :: handle
      type(gorilla_b200_counters_t), intent(out) :: counters
    end function
    integer(c_int) function gorilla_b200_orbit_timestep_events(handle, n, x, vpar, vperp, t_step, boole_initialized, &
                   ind_tetr, iface, t_remain_out, n_pushes, cfg, par_adiab_inv, counter_vpar_0, counter_phi_0, events, &
                   event_cap, n_events) bind(C, name='gorilla_b200_orbit_timestep_events')
      import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t, gorilla_b200_event_settings_t, gorilla_b200_event_t
      type(c_ptr), value        :: handle
      integer(c_int64_t), value :: n, event_cap
      real(c_double)            :: x(3,*), vpar(*), vperp(*), par_adiab_inv(*)
      real(c_double), value     :: t_step
      integer(c_int32_t)        :: boole_initialized(*), ind_tetr(*), iface(*), counter_vpar_0(*), counter_phi_0(*)
      type(c_ptr), value        :: t_remain_out, n_pushes
      type(gorilla_b200_event_settings_t), intent(in) :: cfg
      type(gorilla_b200_event_t) :: events(*)
      integer(c_int64_t), intent(out) :: n_events
    end function
    integer(c_int) function gorilla_b200_invariants(handle, n, x, vpar, vperp, ind_tetr, energy, p_phi, perpinv) &
                   bind(C, name='gorilla_b200_invariants')
      import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t
      type(c_ptr), value        :: handle
      integer(c_int64_t), value :: n
      real(c_double), intent(in):: x(3,*), vpar(*), vperp(*)
      integer(c_int32_t), intent(in) :: ind_tetr(*)
      real(c_double)            :: energy(*), p_phi(*), perpinv(*)
    end function
    integer(c_int) function gorilla_b200_set_host_resort(handle, on) bind(C, name='gorilla_b200_set_host_resort')
      import :: c_int, c_ptr, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: on
    end function
    integer(c_int) function gorilla_b200_set_gather(handle, mode) bind(C, name='gorilla_b200_set_gather')
      import :: c_int, c_ptr, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int32_t), value :: mode
    end function
    integer(c_int) function gorilla_b200_comm_unique_id(id) bind(C, name='gorilla_b200_comm_unique_id')
      import :: c_int, c_char
      character(kind=c_char) :: id(128)
    end function
    integer(c_int) function gorilla_b200_comm_init(handle, id, rank, nranks) bind(C, name='gorilla_b200_comm_init')
      import :: c_int, c_ptr, c_char, c_int32_t
      type(c_ptr), value :: handle
      character(kind=c_char), intent(in) :: id(128)
      integer(c_int32_t), value :: rank, nranks
    end function
    integer(c_int) function gorilla_b200_comm_free(handle) bind(C, name='gorilla_b200_comm_free')
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
    end function
    integer(c_int) function gorilla_b200_shard_range(n_total, rank, nranks, first, count) &
                   bind(C, name='gorilla_b200_shard_range')
      import :: c_int, c_int64_t, c_int32_t
      integer(c_int64_t), value :: n_total
      integer(c_int32_t), value :: rank, nranks
      integer(c_int64_t), intent(out) :: first, count
    end function
    integer(c_int) function gorilla_b200_diag_reset(handle, stream) bind(C, name='gorilla_b200_diag_reset')
      import :: c_int, c_ptr
      type(c_ptr), value :: handle, stream
    end function
    integer(c_int) function gorilla_b200_diag_reduce(handle, n, x, vpar, vperp, ind_tetr, energy_ref, p_phi_ref, &
                   perpinv_ref, diag) bind(C, name='gorilla_b200_diag_reduce')
      import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t, gorilla_b200_diag_t
      type(c_ptr), value        :: handle
      integer(c_int64_t), value :: n
      real(c_double), intent(in):: x(3,*), vpar(*), vperp(*)
      integer(c_int32_t), intent(in) :: ind_tetr(*)
      type(c_ptr), value        :: energy_ref, p_phi_ref, perpinv_ref     ! c_null_ptr: that drift is not formed
      type(gorilla_b200_diag_t), intent(out) :: diag
    end function
    integer(c_int) function gorilla_b200_abi_struct_sizes(sizes) bind(C, name='gorilla_b200_abi_struct_sizes')
      import :: c_int, c_int64_t
      integer(c_int64_t), intent(out) :: sizes(7)
    end function
  end interface

  type(c_ptr), save :: handle = c_null_ptr

contains

  !> Upload what initialize_gorilla left in the module arrays (orbit_timestep_gorilla.f90:151-274).
  subroutine initialize_gorilla_b200(ierr)
    ! coord_system exists in tetra_physics_mod (the value the mesh was built with) and in gorilla_settings_mod (the namelist
    ! entry); both are needed, so the first one is renamed
    use tetra_physics_mod, only: tetra_skew_coord, tetra_physics, cm_over_e, particle_mass, particle_charge, sign_sqg, &
                                 coord_system_mesh => coord_system
    use tetra_grid_mod, only: tetra_grid, ntetr, Rmin, Rmax, Zmin, Zmax
    use tetra_grid_settings_mod, only: grid_kind, grid_size, n_field_periods, sfc_s_min
    use gorilla_settings_mod
    integer, intent(out), optional :: ierr
    type(gorilla_mesh_desc_t) :: md
    type(gorilla_settings_t)  :: st
    type(gorilla_b200_counters_t) :: ct
    type(gorilla_b200_diag_t) :: dg
    type(gorilla_b200_event_t) :: evt
    type(gorilla_b200_event_settings_t) :: evs
    integer(c_int64_t) :: abi(7)
    integer(c_int) :: rc
    md%ntetr = int(ntetr, c_int64_t)
    md%tetra_physics = addr_tetra_physics(tetra_physics)   ! sequence type of 142 doubles -> double[ntetr][142]
    ! the bind(C) types of this module against the layout the library was compiled with
    rc = gorilla_b200_abi_struct_sizes(abi)
    if (rc /= 0 .or. abi(1) /= c_sizeof(st) .or. abi(2) /= c_sizeof(md) .or. abi(3) /= c_sizeof(ct) .or. &
        abi(4) /= c_sizeof(dg) .or. abi(6) /= c_sizeof(evt) .or. abi(7) /= c_sizeof(evs)) then
      print *, 'initialize_gorilla_b200: struct layouts of libgorilla_b200 differ from this module (rebuild both)'
      if (present(ierr)) then
        ierr = 1
        return
      end if
      stop
    end if
    md%tetra_grid    = addr_tetra_grid(tetra_grid)         ! sequence type of 20 integers -> int32[ntetr][20]
    md%cm_over_e = cm_over_e; md%particle_mass = particle_mass; md%particle_charge = particle_charge
    md%sign_sqg = sign_sqg; md%coord_system = coord_system_mesh; md%n_field_periods = n_field_periods
    md%grid_kind = grid_kind; md%grid_size = grid_size; md%pad0 = 0
    md%Rmin = Rmin; md%Rmax = Rmax; md%Zmin = Zmin; md%Zmax = Zmax; md%sfc_s_min = sfc_s_min
    md%tetra_skew_coord = c_null_ptr
    if (handover_processing_kind == 2) md%tetra_skew_coord = addr_tetra_skew(tetra_skew_coord)   ! sequence type, 168 doubles
    st%eps_Phi = eps_Phi; st%coord_system = coord_system; st%ispecies = ispecies
    st%boole_periodic_relocation = merge(1, 0, boole_periodic_relocation)
    st%ipusher = ipusher; st%boole_pusher_ode45 = merge(1, 0, boole_pusher_ode45)
    st%boole_dt_dtau = merge(1, 0, boole_dt_dtau); st%boole_newton_precalc = merge(1, 0, boole_newton_precalc)
    st%poly_order = poly_order; st%i_precomp = i_precomp; st%boole_guess = merge(1, 0, boole_guess)
    st%i_time_tracing_option = i_time_tracing_option; st%handover_processing_kind = handover_processing_kind
    st%boole_adaptive_time_steps = merge(1, 0, boole_adaptive_time_steps)
    st%boole_strong_electric_field = merge(1, 0, boole_strong_electric_field)
    st%boole_grid_for_find_tetra = merge(1, 0, boole_grid_for_find_tetra)
    st%max_n_intermediate_steps = max_n_intermediate_steps; st%desired_delta_energy = desired_delta_energy
    st%rel_err_ode45 = rel_err_ode45
    st%boole_time_Hamiltonian = merge(1, 0, boole_time_Hamiltonian); st%boole_gyrophase = merge(1, 0, boole_gyrophase)
    st%boole_vpar_int = merge(1, 0, boole_vpar_int); st%boole_vpar2_int = merge(1, 0, boole_vpar2_int)
    rc = gorilla_b200_init(md, st, handle)
    if (present(ierr)) then
      ierr = rc
    else if (rc /= 0) then
      print *, 'initialize_gorilla_b200: error code ', rc
      stop
    end if
  end subroutine

  subroutine finalize_gorilla_b200()
    if (c_associated(handle)) call gorilla_b200_free(handle)
    handle = c_null_ptr
  end subroutine

  !> Batched orbit_timestep_gorilla: n independent particles, arrays updated in place.
  subroutine orbit_timestep_gorilla_batch(n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, ierr, &
                                          t_remain_out)
    integer, intent(in)                      :: n
    double precision, intent(inout), target  :: x(3,n), vpar(n), vperp(n)
    double precision, intent(in)             :: t_step
    logical, intent(inout)                   :: boole_initialized(n)
    integer, intent(inout)                   :: ind_tetr(n), iface(n)
    integer, intent(out)                     :: ierr
    double precision, intent(out), optional, target :: t_remain_out(n)
    integer(c_int32_t), allocatable :: binit(:)
    type(c_ptr) :: p_tro
    allocate(binit(n))
    binit = merge(1_c_int32_t, 0_c_int32_t, boole_initialized)
    p_tro = c_null_ptr
    if (present(t_remain_out)) p_tro = c_loc(t_remain_out(1))
    ierr = gorilla_b200_orbit_timestep(handle, int(n, c_int64_t), x, vpar, vperp, t_step, binit, ind_tetr, iface, &
                                       p_tro, c_null_ptr)
    boole_initialized = binit /= 0
  end subroutine

  !> Batch call that also returns pusher_tetra_poly's optional quantities, summed over the pushes of the time step:
  !> optional_quantities(1:4,i) = t_hamiltonian, gyrophase, vpar_int, vpar2_int (type optional_quantities_type,
  !> gorilla_settings_mod.f90:9-15), for the quantities switched on in gorilla.inp.
  subroutine orbit_timestep_gorilla_batch_optional(n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, &
                                                   optional_quantities, ierr)
    integer, intent(in)                      :: n
    double precision, intent(inout), target  :: x(3,n), vpar(n), vperp(n)
    double precision, intent(in)             :: t_step
    logical, intent(inout)                   :: boole_initialized(n)
    integer, intent(inout)                   :: ind_tetr(n), iface(n)
    double precision, intent(out)            :: optional_quantities(4,n)
    integer, intent(out)                     :: ierr
    integer(c_int32_t), allocatable :: binit(:)
    interface
      integer(c_int) function gorilla_b200_orbit_timestep_optional(handle, n, x, vpar, vperp, t_step, boole_initialized, &
                     ind_tetr, iface, t_remain_out, n_pushes, optional_quantities) &
                     bind(C, name='gorilla_b200_orbit_timestep_optional')
        import :: c_int, c_ptr, c_int64_t, c_double, c_int32_t
        type(c_ptr), value :: handle
        integer(c_int64_t), value :: n
        real(c_double) :: x(3,*), vpar(*), vperp(*), optional_quantities(4,*)
        real(c_double), value :: t_step
        integer(c_int32_t) :: boole_initialized(*), ind_tetr(*), iface(*)
        type(c_ptr), value :: t_remain_out, n_pushes
      end function
    end interface
    allocate(binit(n))
    binit = merge(1_c_int32_t, 0_c_int32_t, boole_initialized)
    ierr = gorilla_b200_orbit_timestep_optional(handle, int(n, c_int64_t), x, vpar, vperp, t_step, binit, ind_tetr, &
                                                iface, c_null_ptr, c_null_ptr, optional_quantities)
    boole_initialized = binit /= 0
  end subroutine

  !> The reference's scalar signature (orbit_timestep_gorilla.f90:19), n = 1.
  subroutine orbit_timestep_gorilla(x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out)
    double precision, dimension(3), intent(inout) :: x
    double precision, intent(inout)               :: vpar, vperp
    double precision, intent(in)                  :: t_step
    logical, intent(inout)                        :: boole_initialized
    integer, intent(inout)                        :: ind_tetr, iface
    double precision, intent(out), optional       :: t_remain_out
    double precision :: x1(3,1), vpar1(1), vperp1(1), tro(1)
    logical :: b1(1)
    integer :: it1(1), if1(1), ierr
    x1(:,1) = x; vpar1 = vpar; vperp1 = vperp; b1 = boole_initialized; it1 = ind_tetr; if1 = iface
    call orbit_timestep_gorilla_batch(1, x1, vpar1, vperp1, t_step, b1, it1, if1, ierr, tro)
    if (ierr /= 0) then
      print *, 'orbit_timestep_gorilla (b200): error code ', ierr
      stop
    end if
    x = x1(:,1); vpar = vpar1(1); vperp = vperp1(1); boole_initialized = b1(1); ind_tetr = it1(1); iface = if1(1)
    if (ind_tetr .eq. -1 .and. boole_initialized) print *, 'WARNING: Particle lost.'
    if (present(t_remain_out)) t_remain_out = tro(1)
  end subroutine

  subroutine find_tetra_batch(n, x, vpar, vperp, ind_tetr, iface, sign_t_step, ierr)
    integer, intent(in)             :: n, sign_t_step
    double precision, intent(inout) :: x(3,n)
    double precision, intent(in)    :: vpar(n), vperp(n)
    integer, intent(out)            :: ind_tetr(n), iface(n), ierr
    ierr = gorilla_b200_find_tetra(handle, int(n, c_int64_t), x, vpar, vperp, ind_tetr, iface, int(sign_t_step, c_int32_t))
  end subroutine

  !> Batch call with the event capture of gorilla_plot_orbit_integration (gorilla_plot_mod.f90:585-638): banana tips with
  !> J_par and toroidal mappings go to `events` (event_cap records; n_events may exceed it, the surplus is dropped).
  !> par_adiab_inv / counter_vpar_0 / counter_phi_0 carry the per-particle state between calls (zero them first).
  subroutine orbit_timestep_gorilla_batch_events(n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, cfg, &
                                                 par_adiab_inv, counter_vpar_0, counter_phi_0, events, n_events, ierr)
    integer, intent(in)             :: n
    double precision, intent(inout) :: x(3,n), vpar(n), vperp(n), par_adiab_inv(n)
    double precision, intent(in)    :: t_step
    logical, intent(inout)          :: boole_initialized(n)
    integer, intent(inout)          :: ind_tetr(n), iface(n), counter_vpar_0(n), counter_phi_0(n)
    type(gorilla_b200_event_settings_t), intent(in) :: cfg
    type(gorilla_b200_event_t), intent(out) :: events(:)
    integer(c_int64_t), intent(out) :: n_events
    integer, intent(out)            :: ierr
    integer(c_int32_t), allocatable :: binit(:)
    allocate(binit(n))
    binit = merge(1_c_int32_t, 0_c_int32_t, boole_initialized)
    ierr = gorilla_b200_orbit_timestep_events(handle, int(n, c_int64_t), x, vpar, vperp, t_step, binit, ind_tetr, iface, &
                                              c_null_ptr, c_null_ptr, cfg, par_adiab_inv, counter_vpar_0, counter_phi_0, &
                                              events, int(size(events), c_int64_t), n_events)
    boole_initialized = binit /= 0
  end subroutine

  !> energy_tot_func, p_phi_func and perpinv of every particle (supporting_functions_mod.f90:279-301, 377-408)
  subroutine invariants_b200(n, x, vpar, vperp, ind_tetr, energy, p_phi, perpinv, ierr)
    integer, intent(in)           :: n
    double precision, intent(in)  :: x(3,n), vpar(n), vperp(n)
    integer, intent(in)           :: ind_tetr(n)
    double precision, intent(out) :: energy(n), p_phi(n), perpinv(n)
    integer, intent(out)          :: ierr
    ierr = gorilla_b200_invariants(handle, int(n, c_int64_t), x, vpar, vperp, ind_tetr, energy, p_phi, perpinv)
  end subroutine

  !> .true.: every batch of orbit_timestep_gorilla_batch is sorted by tetrahedron on the device before the push (gather
  !> locality); the caller's particle order is restored before the arrays come back.
  subroutine set_host_resort_b200(on, ierr)
    logical, intent(in)  :: on
    integer, intent(out) :: ierr
    ierr = gorilla_b200_set_host_resort(handle, merge(1_c_int32_t, 0_c_int32_t, on))
  end subroutine

  !> How the order-2 / RK4 kernels gather the tetrahedron records: -1 library default (by mesh size and field content),
  !> 0 per-lane vector loads, 1 per-lane bulk copies, 2 warp-cooperative copies one push ahead.  Results are identical.
  subroutine set_gather_b200(mode, ierr)
    integer, intent(in)  :: mode
    integer, intent(out) :: ierr
    ierr = gorilla_b200_set_gather(handle, int(mode, c_int32_t))
  end subroutine

  !> Multi-GPU: rank 0 creates the id, every rank joins with its own handle (one process per GPU).
  subroutine comm_unique_id_b200(id, ierr)
    character(kind=c_char), intent(out) :: id(128)
    integer, intent(out) :: ierr
    ierr = gorilla_b200_comm_unique_id(id)
  end subroutine
  subroutine comm_init_b200(id, rank, nranks, ierr)
    character(kind=c_char), intent(in) :: id(128)
    integer, intent(in)  :: rank, nranks
    integer, intent(out) :: ierr
    ierr = gorilla_b200_comm_init(handle, id, int(rank, c_int32_t), int(nranks, c_int32_t))
  end subroutine
  subroutine comm_free_b200(ierr)
    integer, intent(out) :: ierr
    ierr = gorilla_b200_comm_free(handle)
  end subroutine
  !> Contiguous shard of n_total particles owned by `rank`: particles first+1 .. first+count (1-based).
  subroutine shard_range_b200(n_total, rank, nranks, first, count)
    integer(c_int64_t), intent(in)  :: n_total
    integer, intent(in)             :: rank, nranks
    integer(c_int64_t), intent(out) :: first, count
    integer(c_int) :: rc
    rc = gorilla_b200_shard_range(n_total, int(rank, c_int32_t), int(nranks, c_int32_t), first, count)
  end subroutine
  subroutine diag_reset_b200(ierr)
    integer, intent(out) :: ierr
    ierr = gorilla_b200_diag_reset(handle, c_null_ptr)
  end subroutine
  !> Counters accumulated since diag_reset_b200 and max / rms drift of energy, perpinv, p_phi against the reference values
  !> (from invariants_b200 at the start), reduced over all ranks of the communicator (collective call).
  subroutine diag_reduce_b200(n, x, vpar, vperp, ind_tetr, energy_ref, p_phi_ref, perpinv_ref, diag, ierr)
    integer, intent(in)                  :: n
    double precision, intent(in)         :: x(3,n), vpar(n), vperp(n)
    integer, intent(in)                  :: ind_tetr(n)
    double precision, intent(in), target :: energy_ref(n), p_phi_ref(n), perpinv_ref(n)
    type(gorilla_b200_diag_t), intent(out) :: diag
    integer, intent(out)                 :: ierr
    ierr = gorilla_b200_diag_reduce(handle, int(n, c_int64_t), x, vpar, vperp, ind_tetr, c_loc(energy_ref(1)), &
                                    c_loc(p_phi_ref(1)), c_loc(perpinv_ref(1)), diag)
  end subroutine

  subroutine get_counters_b200(counters, ierr)
    type(gorilla_b200_counters_t), intent(out) :: counters
    integer, intent(out) :: ierr
    ierr = gorilla_b200_get_counters(handle, counters)
  end subroutine

  ! Addresses of GORILLA's module arrays.  They are declared `allocatable, public, protected` WITHOUT the target attribute
  ! (tetra_physics_mod.f90:85,101; tetra_grid_mod.f90:17), so c_loc() cannot be applied to them directly; they are passed to
  ! an assumed-size dummy that has it.  The arrays are contiguous (no copy-in) and stay allocated until GORILLA deallocates
  ! them, and the library copies them during gorilla_b200_init, so the address is only used inside that call.
  function addr_tetra_physics(a) result(p)
    use tetra_physics_mod, only: tetrahedron_physics
    type(tetrahedron_physics), intent(in), target :: a(*)
    type(c_ptr) :: p
    p = c_loc(a(1))
  end function
  function addr_tetra_grid(a) result(p)
    use tetra_grid_mod, only: tetrahedron_grid
    type(tetrahedron_grid), intent(in), target :: a(*)
    type(c_ptr) :: p
    p = c_loc(a(1))
  end function
  function addr_tetra_skew(a) result(p)
    use tetra_physics_mod, only: tetrahedron_skew_coord
    type(tetrahedron_skew_coord), intent(in), target :: a(*)
    type(c_ptr) :: p
    p = c_loc(a(1))
  end function

end module orbit_timestep_gorilla_b200_mod
