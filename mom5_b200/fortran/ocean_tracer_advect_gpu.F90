! ocean_tracer_advect_gpu.F90 -- ISO_C_BINDING shim between MOM5's ocean_tracer_advect_mod and libmom5adv.so.
!
! SOURCE ONLY: this image has no Fortran compiler, so the shim is shipped uncompiled (see INTEGRATION.md).
! It is deliberately thin: it unpacks the derived types into raw arrays and calls the C ABI of
! include/mom5adv.h.  Everything it calls is exercised from Python/ctypes by tests/test_gpu_parity.py.
!
! Call sites it replaces (OTA = src/mom5/ocean_tracers/ocean_tracer_advect.F90):
!   OTA:2078-2080   call advect_tracer_sweby_all(...)                 -> gpu_advect_tracer_sweby_all
!   OTA:1935-1962   the upwind / quicker / mdfl_sweby / dst_linear arms of horz_advect_tracer -> gpu_horz_advect_tracer
!   OTA:2126-2140   the upwind / quicker arms of vert_advect_tracer   -> gpu_vert_advect_tracer
!   OTA:507-756     ocean_tracer_advect_init (after mdfl_init / quicker_init) -> gpu_tracer_advect_init
module ocean_tracer_advect_gpu_mod
  use, intrinsic :: iso_c_binding
  use ocean_types_mod,  only: ocean_grid_type, ocean_domain_type, ocean_prog_tracer_type, &
                              ocean_adv_vel_type, ocean_thickness_type, ocean_time_type
  use mpp_mod,          only: mpp_error, FATAL, mpp_pe, mpp_npes, mpp_root_pe, mpp_broadcast
  use ocean_parameters_mod, only: ADVECT_MDPPM
  implicit none
  private
  public :: gpu_tracer_advect_init, gpu_advect_tracer_sweby_all, gpu_horz_advect_tracer, gpu_vert_advect_tracer, &
            gpu_compute_adv_diss, gpu_tracer_advect_end

  type, bind(C) :: mom5adv_grid
     integer(c_int) :: isc, iec, jsc, jec, nk
     integer(c_int) :: ni_global, nj_global, layout_x, layout_y
     type(c_ptr)    :: x_extent, y_extent
     integer(c_int) :: cyclic_x, cyclic_y, tripolar, have_obc
     type(c_ptr)    :: dat, datr, dxt, dyt, dxte, dyte, dxtn, dytn, dzt, tmask
  end type mom5adv_grid

  interface
     function mom5adv_last_error() bind(C, name='mom5adv_last_error') result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function
     function mom5adv_set_device(node_local_rank) bind(C, name='mom5adv_set_device') result(rc)
       import :: c_int
       integer(c_int), value :: node_local_rank
       integer(c_int) :: rc
     end function
     function mom5adv_comm_unique_id(id) bind(C, name='mom5adv_comm_unique_id') result(rc)
       import :: c_int, c_char
       character(kind=c_char) :: id(128)
       integer(c_int) :: rc
     end function
     function mom5adv_comm_create(id, rank, nranks, comm) bind(C, name='mom5adv_comm_create') result(rc)
       import :: c_int, c_char, c_ptr
       character(kind=c_char) :: id(128)
       integer(c_int), value :: rank, nranks
       type(c_ptr) :: comm
       integer(c_int) :: rc
     end function
     function mom5adv_init(grid, ntracers_max, comm, handle) bind(C, name='mom5adv_init') result(rc)
       import :: c_int, c_ptr, mom5adv_grid
       type(mom5adv_grid) :: grid
       integer(c_int), value :: ntracers_max
       type(c_ptr), value :: comm
       type(c_ptr) :: handle
       integer(c_int) :: rc
     end function
     function mom5adv_finalize(handle) bind(C, name='mom5adv_finalize') result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int) :: rc
     end function
     function mom5adv_sweby_all(handle, ntr, dtime, T, th, adv, u, v, w, rho, fx, fy, fz, ax, ay, az) &
          bind(C, name='mom5adv_sweby_all') result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       integer(c_int), value :: ntr
       real(c_double), value :: dtime
       type(c_ptr) :: T(*), th(*), adv(*)
       type(c_ptr), value :: u, v, w, rho
       type(c_ptr), value :: fx, fy, fz, ax, ay, az     ! arrays of ntr pointers or c_null_ptr
       integer(c_int) :: rc
     end function
     function mom5adv_horz(handle, scheme, dtime, Tm1, Tt, tlimit, limit_with_upwind, u, v, w, rho, th, wrk1, fx, fy, fz) &
          bind(C, name='mom5adv_horz') result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: handle
       integer(c_int), value :: scheme, limit_with_upwind
       real(c_double), value :: dtime
       type(c_ptr), value :: Tm1, Tt, tlimit, u, v, w, rho, th, wrk1, fx, fy, fz
       integer(c_int) :: rc
     end function
     function mom5adv_adv_diss(handle, horz_scheme, vert_scheme, dtime, conversion, T_tau, tlimit, limit_with_upwind, &
          u, v, w, rho_tau, rho_taup1, advect_tendency, adv_diss, t2_tendency) bind(C, name='mom5adv_adv_diss') result(rc)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value    :: handle
       integer(c_int), value :: horz_scheme, vert_scheme, limit_with_upwind
       real(c_double), value :: dtime, conversion
       type(c_ptr), value    :: T_tau, tlimit, u, v, w, rho_tau, rho_taup1, advect_tendency, adv_diss, t2_tendency
       integer(c_int)        :: rc
     end function mom5adv_adv_diss
     function mom5adv_set_ppm_limiters(handle, ppm_hlimiter, ppm_vlimiter) bind(C, name='mom5adv_set_ppm_limiters') result(rc)
       import :: c_ptr, c_int
       type(c_ptr), value    :: handle
       integer(c_int), value :: ppm_hlimiter, ppm_vlimiter
       integer(c_int)        :: rc
     end function mom5adv_set_ppm_limiters
     function mom5adv_vert(handle, scheme, Tm1, Tt, tlimit, w, th, wrk1, fz) bind(C, name='mom5adv_vert') result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: handle
       integer(c_int), value :: scheme
       type(c_ptr), value :: Tm1, Tt, tlimit, w, th, wrk1, fz
       integer(c_int) :: rc
     end function
  end interface

  type(c_ptr), save :: handle = c_null_ptr, comm = c_null_ptr

contains

  subroutine check(rc, where)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: cmsg(:)
    character(len=512) :: msg
    integer :: n
    if (rc == 0) return
    call c_f_pointer(mom5adv_last_error(), cmsg, [512])
    msg = ' '
    do n = 1, 512
       if (cmsg(n) == c_null_char) exit
       msg(n:n) = cmsg(n)
    end do
    call mpp_error(FATAL, '==>Error from ocean_tracer_advect_gpu_mod ('//trim(where)//'): '//trim(msg))
  end subroutine check

  ! called at the end of ocean_tracer_advect_init (OTA:507-756)
  subroutine gpu_tracer_advect_init(Grid, Domain, nk, num_prog_tracers, have_obc)
    type(ocean_grid_type),   intent(in), target :: Grid
    type(ocean_domain_type), intent(in)         :: Domain
    integer, intent(in) :: nk, num_prog_tracers
    logical, intent(in) :: have_obc
    type(mom5adv_grid) :: g
    character(kind=c_char) :: id(128)
    integer :: id_int(128)
    integer :: n
    ! one rank drives one GPU: bind this rank to device (node-local rank mod device count) BEFORE any other call -- otherwise every
    ! rank of a node lands on device 0 and ncclCommInitRank refuses the duplicate GPU.  Ranks of a node are consecutive PEs in FMS
    ! (mpp_pe() - mpp_root_pe() counts from 0), so the global rank serves as the node-local one modulo the device count.
    call check(mom5adv_set_device(int(mpp_pe() - mpp_root_pe(), c_int)), 'set_device')
    g%isc = Domain%isc; g%iec = Domain%iec; g%jsc = Domain%jsc; g%jec = Domain%jec; g%nk = nk
    g%ni_global = Grid%ni; g%nj_global = Grid%nj
    g%layout_x = Domain%layout(1); g%layout_y = Domain%layout(2)
    g%x_extent = c_null_ptr; g%y_extent = c_null_ptr          ! mpp_compute_extent rule (no user extents)
    g%cyclic_x = merge(1, 0, Grid%cyclic_x); g%cyclic_y = merge(1, 0, Grid%cyclic_y)
    g%tripolar = merge(1, 0, Grid%tripolar);  g%have_obc = merge(1, 0, have_obc)
    g%dat  = c_loc(Grid%dat);  g%datr = c_loc(Grid%datr); g%dxt  = c_loc(Grid%dxt);  g%dyt  = c_loc(Grid%dyt)
    g%dxte = c_loc(Grid%dxte); g%dyte = c_loc(Grid%dyte); g%dxtn = c_loc(Grid%dxtn); g%dytn = c_loc(Grid%dytn)
    g%dzt  = c_loc(Grid%dzt);  g%tmask = c_loc(Grid%tmask)
    if (mpp_npes() > 1) then
       ! bootstrap the NCCL communicator over the model's own MPI: rank 0 makes the id, mpp_broadcast ships it.
       ! Requires the FMS pelist order pe = ix + layout_x*iy, which is what mpp_define_domains produces.
       if (mpp_pe() == mpp_root_pe()) call check(mom5adv_comm_unique_id(id), 'comm_unique_id')
       ! mpp_broadcast has no scalar-character specific (mpp_broadcast_char wants data(:), length, from_pe): ship the 128 bytes as
       ! an integer array through the rank-1 integer specific mpp_broadcast(data(:), length, from_pe)
       do n = 1, 128
          id_int(n) = ichar(id(n))
       end do
       call mpp_broadcast(id_int, 128, mpp_root_pe())
       do n = 1, 128
          id(n) = char(id_int(n), kind=c_char)     ! char, not achar: the id bytes use all 256 values
       end do
       call check(mom5adv_comm_create(id, int(mpp_pe() - mpp_root_pe(), c_int), int(mpp_npes(), c_int), comm), 'comm_create')
    end if
    call check(mom5adv_init(g, int(num_prog_tracers, c_int), comm, handle), 'init')
  end subroutine gpu_tracer_advect_init

  ! replaces `call advect_tracer_sweby_all(Time, Adv_vel, Dens, T_prog, Thickness, dtime)` (OTA:2078-2080)
  subroutine gpu_advect_tracer_sweby_all(Time, Adv_vel, T_prog, Thickness, dtime)
    type(ocean_time_type),        intent(in)            :: Time
    type(ocean_adv_vel_type),     intent(in),    target :: Adv_vel
    type(ocean_prog_tracer_type), intent(inout), target :: T_prog(:)
    type(ocean_thickness_type),   intent(in),    target :: Thickness
    real,                         intent(in)            :: dtime
    type(c_ptr) :: pT(size(T_prog)), pth(size(T_prog)), padv(size(T_prog))
    integer :: n, isd, jsd
    isd = lbound(T_prog(1)%field, 1); jsd = lbound(T_prog(1)%field, 2)
    do n = 1, size(T_prog)
       pT(n)   = c_loc(T_prog(n)%field(isd, jsd, 1, Time%taum1))
       pth(n)  = c_loc(T_prog(n)%th_tendency)
       padv(n) = c_loc(T_prog(n)%wrk1)
    end do
    call check(mom5adv_sweby_all(handle, int(size(T_prog), c_int), real(dtime, c_double), pT, pth, padv, &
         c_loc(Adv_vel%uhrho_et), c_loc(Adv_vel%vhrho_nt), c_loc(Adv_vel%wrho_bt), &
         c_loc(Thickness%rho_dzt(isd, jsd, 1, Time%tau)), &
         c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr, c_null_ptr), 'sweby_all')
    ! diagnostics (OTA:4223-4229 etc.): pass arrays of pointers instead of c_null_ptr for the registered ids and
    ! hand the returned fields to diagnose_3d exactly as the reference does.
  end subroutine gpu_advect_tracer_sweby_all

  ! replaces one arm of the select case in horz_advect_tracer (OTA:1933-1996) incl. th_tendency += wrk1
  subroutine gpu_horz_advect_tracer(Time, Adv_vel, Thickness, Tracer, dtime, limit_with_upwind, flux_x, flux_y, flux_z)
    type(ocean_time_type),        intent(in)            :: Time
    type(ocean_adv_vel_type),     intent(in),    target :: Adv_vel
    type(ocean_thickness_type),   intent(in),    target :: Thickness
    type(ocean_prog_tracer_type), intent(inout), target :: Tracer
    real,                         intent(in)            :: dtime
    logical,                      intent(in)            :: limit_with_upwind
    real, dimension(:,:,:),       intent(inout), target :: flux_x, flux_y, flux_z   ! the module work arrays
    integer :: isd, jsd
    isd = lbound(Tracer%field, 1); jsd = lbound(Tracer%field, 2)
    if (Tracer%horz_advect_scheme == ADVECT_MDPPM) then   ! per-tracer field-table entries (ocean_tracer.F90:1006,1051)
       call check(mom5adv_set_ppm_limiters(handle, int(Tracer%ppm_hlimiter, c_int), int(Tracer%ppm_vlimiter, c_int)), 'ppm limiters')
    endif
    call check(mom5adv_horz(handle, int(Tracer%horz_advect_scheme, c_int), real(dtime, c_double), &
         c_loc(Tracer%field(isd, jsd, 1, Time%taum1)), c_loc(Tracer%field(isd, jsd, 1, Time%tau)), &
         c_loc(Tracer%tmask_limit), merge(1_c_int, 0_c_int, limit_with_upwind), &
         c_loc(Adv_vel%uhrho_et), c_loc(Adv_vel%vhrho_nt), c_loc(Adv_vel%wrho_bt), &
         c_loc(Thickness%rho_dzt(isd, jsd, 1, Time%tau)), c_loc(Tracer%th_tendency), c_loc(Tracer%wrk1), &
         c_loc(flux_x), c_loc(flux_y), c_loc(flux_z)), 'horz_advect_tracer')
  end subroutine gpu_horz_advect_tracer

  ! replaces one arm of the select case in vert_advect_tracer (OTA:2124-2168)
  subroutine gpu_vert_advect_tracer(Time, Adv_vel, Tracer, flux_z)
    type(ocean_time_type),        intent(in)            :: Time
    type(ocean_adv_vel_type),     intent(in),    target :: Adv_vel
    type(ocean_prog_tracer_type), intent(inout), target :: Tracer
    real, dimension(:,:,:),       intent(inout), target :: flux_z
    integer :: isd, jsd
    isd = lbound(Tracer%field, 1); jsd = lbound(Tracer%field, 2)
    call check(mom5adv_vert(handle, int(Tracer%vert_advect_scheme, c_int), &
         c_loc(Tracer%field(isd, jsd, 1, Time%taum1)), c_loc(Tracer%field(isd, jsd, 1, Time%tau)), &
         c_loc(Tracer%tmask_limit), c_loc(Adv_vel%wrho_bt), c_loc(Tracer%th_tendency), c_loc(Tracer%wrk1), &
         c_loc(flux_z)), 'vert_advect_tracer')
  end subroutine gpu_vert_advect_tracer

  ! replaces compute_adv_diss (OTA:7547-7712) up to its diagnose_3d calls: wrk4 = adv_diss, wrk1 = tendency of the squared tracer
  subroutine gpu_compute_adv_diss(Time, Adv_vel, Thickness, Tracer, dtime, limit_with_upwind, advect_tendency, wrk1, wrk4)
    type(ocean_time_type),        intent(in)            :: Time
    type(ocean_adv_vel_type),     intent(in),    target :: Adv_vel
    type(ocean_thickness_type),   intent(in),    target :: Thickness
    type(ocean_prog_tracer_type), intent(inout), target :: Tracer
    real,                         intent(in)            :: dtime
    logical,                      intent(in)            :: limit_with_upwind
    real, dimension(:,:,:),       intent(in),    target :: advect_tendency
    real, dimension(:,:,:),       intent(inout), target :: wrk1, wrk4
    integer :: isd, jsd
    isd = lbound(Tracer%field, 1); jsd = lbound(Tracer%field, 2)
    if (Tracer%horz_advect_scheme == ADVECT_MDPPM) then
       call check(mom5adv_set_ppm_limiters(handle, int(Tracer%ppm_hlimiter, c_int), int(Tracer%ppm_vlimiter, c_int)), 'ppm limiters')
    endif
    call check(mom5adv_adv_diss(handle, int(Tracer%horz_advect_scheme, c_int), int(Tracer%vert_advect_scheme, c_int), &
         real(dtime, c_double), real(Tracer%conversion, c_double), c_loc(Tracer%field(isd, jsd, 1, Time%tau)), &
         c_loc(Tracer%tmask_limit), merge(1_c_int, 0_c_int, limit_with_upwind), &
         c_loc(Adv_vel%uhrho_et), c_loc(Adv_vel%vhrho_nt), c_loc(Adv_vel%wrho_bt), &
         c_loc(Thickness%rho_dzt(isd, jsd, 1, Time%tau)), c_loc(Thickness%rho_dzt(isd, jsd, 1, Time%taup1)), &
         c_loc(advect_tendency), c_loc(wrk4), c_loc(wrk1)), 'compute_adv_diss')
  end subroutine gpu_compute_adv_diss

  subroutine gpu_tracer_advect_end()
    integer(c_int) :: rc
    rc = mom5adv_finalize(handle)
    handle = c_null_ptr
  end subroutine gpu_tracer_advect_end

end module ocean_tracer_advect_gpu_mod
