! mcfost_b200_shim.f90 -- the ISO_C_BINDING layer a maintainer adds to MCFOST (src/) to run the Monte Carlo photon-packet
! loop on B200 GPUs through libmcfost_b200.so (include/mcfost_b200.h).
!
! It replaces the BODY of  subroutine mc_photon_loop(lambda_in, p_lambda_in, n_photons2, n_phot_lim, nnfot1_start,
! laffichage)  (dust_transfer.f90:439-572): same dummy arguments, same module variables read and written.  Every member of
! every ABI struct is assigned here, in the order of the header; shim/c_driver.c is the same sequence in C and is compiled
! and run by the test-suite (tests/test_c_driver.py).  THIS FILE HAS NOT SEEN A FORTRAN COMPILER in the environment it was
! written in (none is installed); module / variable names are those of cpinte/mcfost 4.1.13.
!
! Conventions follow the reference's own C boundary voro_C (Voronoi.f90:70-96 <-> voro++_wrapper.cpp:41-44): bind(C), value
! scalars, caller-allocated arrays passed with c_loc, a non-zero return code -> call error() -> exit(1) (messages.f90:27-46).
module mcfost_b200_shim
  use iso_c_binding
  use mcfost_env, only : dp
  use parameters
  use constants
  use messages, only : error
  use grid
  use cylindrical_grid
  use Voronoi_grid
  use dust_prop
  use grains
  use density, only : dust_density_o_n_grains => densite_pouss_o_n_grains   ! (n_grains|n_zones, n_cells), density.f90:32
  use thermal_emission
  use Temperature, only : tab_Temp, T_min
  use radiation_field
  use stars
  use wavelengths, only : n_lambda, tab_lambda, tab_delta_lambda
  use dust_ray_tracing
  use output
  use naleat, only : seed
  implicit none

  ! ---------------------------------------------------------------- the structs of include/mcfost_b200.h, member for member
  type, bind(C) :: mcb_grid
     integer(c_int32_t) :: kind, l3D, n_rad, nz, n_az, n_cells
     real(c_double)     :: Rmax2, zmaxmax
     type(c_ptr) :: r_lim, r_lim_2, r_lim_3, z_lim, zmax, tan_theta_lim, theta_lim, tan_phi_lim, volume
     integer(c_int32_t) :: n_cells_tot
     type(c_ptr) :: cell_map_i, cell_map_j, cell_map_k
     type(c_ptr) :: vor_xyz, vor_h, vor_first, vor_last, vor_was_cut, vor_is_star, vor_is_star_neighbour, neighbours_list
     integer(c_int64_t) :: n_neighbours_tot
     real(c_float)  :: wall_x(4,6)
     real(c_double) :: cutting_distance_o_h
     integer(c_int32_t) :: n_stars
     type(c_ptr) :: star_xyzr, star_icell, star_out_model
     type(c_ptr) :: w_lim, sin_phi_lim, cos_phi_lim
  end type mcb_grid

  type, bind(C) :: mcb_opacity
     integer(c_int32_t) :: n_lambda, p_n_cells, p_n_lambda_pos, n_T
     type(c_ptr) :: kappa, kappa_abs_LTE, kappa_factor, tab_albedo_pos, tab_g_pos
     type(c_ptr) :: prob_s11_pos, tab_s11_pos, tab_s12_o_s11_pos, tab_s22_o_s11_pos, tab_s33_o_s11_pos, tab_s34_o_s11_pos, tab_s44_o_s11_pos
     type(c_ptr) :: log_Qcool_minus_extra_heating, kdB_dT_CDF, tab_Temp
     real(c_float) :: T_min
  end type mcb_opacity

  type, bind(C) :: mcb_grains
     integer(c_int32_t) :: n_grains_tot, n_dens
     integer(c_int32_t) :: grain_RE_LTE_start, grain_RE_LTE_end, grain_RE_nLTE_start, grain_RE_nLTE_end, grain_nRE_start, grain_nRE_end
     type(c_ptr) :: grain_zone, n_grains, dust_density_o_n_grains
     type(c_ptr) :: C_abs, C_abs_norm, C_sca, tab_g
     type(c_ptr) :: prob_s11, tab_s11, tab_s12, tab_s22, tab_s33, tab_s34, tab_s44
     type(c_ptr) :: ksca_CDF
     type(c_ptr) :: kappa_abs_nLTE, kabs_nLTE_CDF, log_E_em_1grain, kdB_dT_1grain_nLTE_CDF
     type(c_ptr) :: kappa_abs_RE, proba_abs_RE, Proba_abs_RE_LTE, Proba_abs_RE_LTE_p_nLTE
     type(c_ptr) :: log_E_em_1grain_nRE, kdB_dT_1grain_nRE_CDF, l_RE, J0
     type(c_ptr) :: kdB_dT_1grain_LTE_CDF
  end type mcb_grains

  type, bind(C) :: mcb_emission
     type(c_ptr) :: spectre_emission_cumul, frac_E_stars, frac_E_disk, prob_E_cell, CDF_E_star
     real(c_double) :: L_packet_th, E_paquet, R_ISM, centre_ISM(3)
     type(c_ptr) :: correct_E_emission
  end type mcb_emission

  type, bind(C) :: mcb_run_params
     integer(c_int32_t) :: lambda_in, p_lambda_in, n_photons2
     real(c_float)      :: n_phot_lim
     integer(c_int32_t) :: nnfot1_start, laffichage
     integer(c_int32_t) :: n_photons_loop
     integer(c_int32_t) :: letape_th, lmono, lmono0
     integer(c_int32_t) :: lscatt_ray_tracing1, lscatt_ray_tracing2
     integer(c_int32_t) :: lsepar_pola, lsepar_contrib
     integer(c_int32_t) :: lscattering_method1, lmethod_aniso1, lisotropic
     integer(c_int32_t) :: l_sym_centrale, l_sym_axiale
     integer(c_int32_t) :: lonly_LTE, lxJ_abs_step1, lxJ_abs
     integer(c_int32_t) :: N_thet, N_phi, capt_sup
     integer(c_int32_t) :: RT_n_incl, RT_n_az
     type(c_ptr) :: tab_u_rt, tab_v_rt, tab_w_rt
     integer(c_int64_t) :: seed
     integer(c_int32_t) :: call_index
     integer(c_int32_t) :: rank, n_ranks, reset_tallies
     integer(c_int32_t) :: loutput_mc, n_theta_I, n_phi_I
     integer(c_int32_t) :: lonly_nLTE, lRE_nLTE, lnRE
     integer(c_int32_t) :: low_mem_th_emission_nLTE, low_mem_scattering
     integer(c_int32_t) :: npix_x, npix_y
     real(c_float)      :: zoom
     real(c_double)     :: map_size, cos_disk, sin_disk
     integer(c_int32_t) :: l_sym_ima, lonly_capt_interet, capt_inf, lorigine, capt_interet
     integer(c_int32_t) :: low_mem_th_emission, lweight_emission, lspot
     real(c_float)      :: T_spot, surf_fraction_spot, theta_spot, phi_spot
     real(c_double)     :: star1_T
     type(c_ptr)        :: tab_lambda
     integer(c_int32_t) :: lxN_abs
     integer(c_int32_t) :: lMRW
     real(c_float)      :: gamma_MRW
     integer(c_int32_t) :: lcount_sent
     real(c_float)      :: max_inflight_fraction
     integer(c_int32_t) :: lISM_loop
  end type mcb_run_params

  type, bind(C) :: mcb_tallies
     type(c_ptr) :: xKJ_abs, xJ_abs, xT_ech, n_phot_envoyes
     type(c_ptr) :: sed, sed_q, sed_u, sed_v, n_phot_sed, sed_star, sed_star_scat, sed_disk, sed_disk_scat
     type(c_ptr) :: xI_scatt
     integer(c_int32_t) :: N_type_flux
     type(c_ptr) :: I_spec, I_spec_star
     type(c_ptr) :: stats
     type(c_ptr) :: xT_ech_1grain, xT_ech_1grain_nRE, E_abs_nRE
     type(c_ptr) :: stokes_map, star_origin, disk_origin, xN_abs
  end type mcb_tallies

  interface
     integer(c_int) function mcfost_b200_multi_init(n_gpus, devices, m) bind(C, name='mcfost_b200_multi_init')
       import ; integer(c_int), value :: n_gpus ; type(c_ptr), value :: devices ; type(c_ptr) :: m
     end function
     subroutine mcfost_b200_multi_finalize(m) bind(C, name='mcfost_b200_multi_finalize')
       import ; type(c_ptr), value :: m
     end subroutine
     integer(c_int) function mcfost_b200_multi_upload_grid(m, g) bind(C, name='mcfost_b200_multi_upload_grid')
       import ; type(c_ptr), value :: m ; type(mcb_grid) :: g
     end function
     integer(c_int) function mcfost_b200_multi_upload_dark_zone(m, dz) bind(C, name='mcfost_b200_multi_upload_dark_zone')
       import ; type(c_ptr), value :: m ; integer(c_int32_t) :: dz(*)
     end function
     integer(c_int) function mcfost_b200_multi_upload_opacity(m, o) bind(C, name='mcfost_b200_multi_upload_opacity')
       import ; type(c_ptr), value :: m ; type(mcb_opacity) :: o
     end function
     integer(c_int) function mcfost_b200_multi_upload_emission(m, e) bind(C, name='mcfost_b200_multi_upload_emission')
       import ; type(c_ptr), value :: m ; type(mcb_emission) :: e
     end function
     integer(c_int) function mcfost_b200_multi_upload_grains(m, g) bind(C, name='mcfost_b200_multi_upload_grains')
       import ; type(c_ptr), value :: m ; type(mcb_grains) :: g
     end function
     integer(c_int) function mcfost_b200_multi_run(m, r, t) bind(C, name='mcfost_b200_multi_run')
       import ; type(c_ptr), value :: m ; type(mcb_run_params) :: r ; type(mcb_tallies) :: t
     end function
     type(c_ptr) function mcfost_b200_multi_handle(m, i) bind(C, name='mcfost_b200_multi_handle')
       import ; type(c_ptr), value :: m ; integer(c_int), value :: i
     end function
     integer(c_int) function mcfost_b200_define_dark_zone(h, lambda, tau_max, r_grid, z_grid, n_regions, iRmin, iRmax, dust_sum, &
          l_dark, ri_in, ri_out, zj_sup, zj_inf, l_is_dark) bind(C, name='mcfost_b200_define_dark_zone')
       import
       type(c_ptr), value :: h
       integer(c_int32_t), value :: lambda, n_regions
       real(c_float), value :: tau_max
       real(c_double), intent(in) :: r_grid(*), z_grid(*)
       integer(c_int32_t), intent(in) :: iRmin(*), iRmax(*)
       type(c_ptr), value :: dust_sum
       integer(c_int32_t) :: l_dark(*), ri_in(*), ri_out(*), zj_sup(*), zj_inf(*)
       integer(c_int32_t) :: l_is_dark
     end function
     integer(c_int) function mcfost_b200_multi_n_gpus(m) bind(C, name='mcfost_b200_multi_n_gpus')
       import ; type(c_ptr), value :: m
     end function
     integer(c_int) function mcfost_b200_init_reemission(h, tab_lambda, tab_delta_lambda, logQ, cdf) bind(C, name='mcfost_b200_init_reemission')
       import
       type(c_ptr), value :: h, logQ, cdf
       real(c_double), intent(in) :: tab_lambda(*), tab_delta_lambda(*)
     end function
     integer(c_int) function mcfost_b200_repartition_energie(h, lambda_first, lambda_last, Tdust, tab_lambda, E_stars, E_ISM, weight, &
          E_disk, frac_E_stars, frac_E_disk, weight_norm, prob_E_cell) bind(C, name='mcfost_b200_repartition_energie')
       import
       type(c_ptr), value :: h, weight, weight_norm, prob_E_cell
       integer(c_int32_t), value :: lambda_first, lambda_last
       real(c_float), intent(in) :: Tdust(*)
       real(c_double), intent(in) :: tab_lambda(*), E_stars(*), E_ISM(*)
       real(c_double) :: E_disk(*), frac_E_stars(*), frac_E_disk(*)
     end function
     integer(c_int) function mcfost_b200_index_cell(h, n, x, y, z, icell) bind(C, name='mcfost_b200_index_cell')
       import
       type(c_ptr), value :: h
       integer(c_int64_t), value :: n
       real(c_double), intent(in) :: x(*), y(*), z(*)
       integer(c_int32_t) :: icell(*)
     end function
     integer(c_int) function mcfost_b200_optical_length_tot(h, n, lambda, x, y, z, u, v, w, icell, tau_tot, lmin, lmax, n_steps) &
          bind(C, name='mcfost_b200_optical_length_tot')
       import
       type(c_ptr), value :: h, n_steps
       integer(c_int64_t), value :: n
       integer(c_int32_t), value :: lambda
       real(c_double), intent(in) :: x(*), y(*), z(*), u(*), v(*), w(*)
       integer(c_int32_t), intent(in) :: icell(*)
       real(c_double) :: tau_tot(*), lmin(*), lmax(*)
     end function
     integer(c_int) function mcfost_b200_compute_column(h, lambda, factor, cx, cy, cz, column) bind(C, name='mcfost_b200_compute_column')
       import
       type(c_ptr), value :: h, factor
       integer(c_int32_t), value :: lambda
       real(c_double), intent(in) :: cx(*), cy(*), cz(*)
       real(c_float) :: column(*)
     end function
  end interface

  type(c_ptr), save :: b200 = c_null_ptr          ! the multi-GPU object (n_gpus = 1 is the single-GPU case)
  integer, save :: b200_call_index = 0
  ! contiguous copies of what MCFOST keeps in derived types or as `logical`
  real(c_double), allocatable, target, save :: star_xyzr(:,:)
  integer(c_int32_t), allocatable, target, save :: star_icell(:), star_out(:), dark_i32(:), l_RE_i32(:,:), grain_zone_i32(:)
  real(c_double), allocatable, target, save :: vor_xyz(:,:), vor_h(:)
  integer(c_int32_t), allocatable, target, save :: vor_first(:), vor_last(:), vor_cut(:), vor_star(:), vor_star_nb(:)
  real(c_double), target, save :: stats(12), E_abs_nRE_c

contains

  subroutine b200_check(ierr, what)
    integer(c_int), intent(in) :: ierr ; character(len=*), intent(in) :: what
    if (ierr /= 0) call error("mcfost_b200: "//what//" failed")      ! messages.f90:27-46 -> exit(1); there is no CPU fallback
  end subroutine b200_check

  integer(c_int32_t) function l2i(l)
    logical, intent(in) :: l
    l2i = merge(1_c_int32_t, 0_c_int32_t, l)
  end function l2i

  !------------------------------------------------------------------------------------------------------------------
  ! once, at the end of init_dust_transfer (dust_transfer.f90:41-340): grid, stars, opacity and thermal tables
  subroutine b200_upload_model(n_gpus)
    integer, intent(in) :: n_gpus
    type(mcb_grid) :: g ; type(mcb_opacity) :: o
    integer :: i

    call b200_check(mcfost_b200_multi_init(int(n_gpus, c_int), c_null_ptr, b200), "multi_init")

    if (allocated(star_xyzr)) deallocate(star_xyzr, star_icell, star_out)
    allocate(star_xyzr(4, n_stars), star_icell(n_stars), star_out(n_stars))
    do i = 1, n_stars
       star_xyzr(:, i) = (/ star(i)%x, star(i)%y, star(i)%z, star(i)%r /)
       star_icell(i) = star(i)%icell ; star_out(i) = l2i(star(i)%out_model)
    enddo

    g%kind = merge(3, merge(2, 1, lspherical), lVoronoi) ; g%l3D = l2i(l3D)
    g%n_rad = n_rad ; g%nz = nz ; g%n_az = n_az
    g%n_cells = n_cells
    g%Rmax2 = Rmax2 ; g%zmaxmax = zmaxmax
    g%r_lim = c_null_ptr ; g%r_lim_2 = c_null_ptr ; g%r_lim_3 = c_null_ptr ; g%z_lim = c_null_ptr ; g%zmax = c_null_ptr
    g%tan_theta_lim = c_null_ptr ; g%theta_lim = c_null_ptr ; g%tan_phi_lim = c_null_ptr
    g%n_cells_tot = 0 ; g%cell_map_i = c_null_ptr ; g%cell_map_j = c_null_ptr ; g%cell_map_k = c_null_ptr
    g%vor_xyz = c_null_ptr ; g%vor_h = c_null_ptr ; g%vor_first = c_null_ptr ; g%vor_last = c_null_ptr
    g%vor_was_cut = c_null_ptr ; g%vor_is_star = c_null_ptr ; g%vor_is_star_neighbour = c_null_ptr ; g%neighbours_list = c_null_ptr
    g%n_neighbours_tot = 0 ; g%wall_x = 0.0 ; g%cutting_distance_o_h = 0.0_dp
    g%w_lim = c_null_ptr ; g%sin_phi_lim = c_null_ptr ; g%cos_phi_lim = c_null_ptr
    if (.not.lVoronoi) then
       g%r_lim = c_loc(r_lim) ; g%r_lim_2 = c_loc(r_lim_2) ; g%r_lim_3 = c_loc(r_lim_3)           ! (0:n_rad)
       if (lcylindrical) then
          g%z_lim = c_loc(z_lim) ; g%zmax = c_loc(zmax)                                             ! (n_rad, nz+2), (n_rad)
       else
          g%tan_theta_lim = c_loc(tan_theta_lim) ; g%theta_lim = c_loc(theta_lim) ; g%w_lim = c_loc(w_lim)   ! (0:nz)
       endif
       if (l3D) then
          g%tan_phi_lim = c_loc(tan_phi_lim) ; g%sin_phi_lim = c_loc(sin_phi_lim) ; g%cos_phi_lim = c_loc(cos_phi_lim)   ! (n_az)
       endif
       g%n_cells_tot = size(cell_map_i)            ! lets the library verify its closed-form numbering (MCB_ERR_CELL_MAP)
       g%cell_map_i = c_loc(cell_map_i) ; g%cell_map_j = c_loc(cell_map_j) ; g%cell_map_k = c_loc(cell_map_k)
    else
       if (allocated(vor_xyz)) deallocate(vor_xyz, vor_h, vor_first, vor_last, vor_cut, vor_star, vor_star_nb)
       allocate(vor_xyz(3, n_cells), vor_h(n_cells), vor_first(n_cells), vor_last(n_cells), vor_cut(n_cells), vor_star(n_cells), vor_star_nb(n_cells))
       do i = 1, n_cells
          vor_xyz(:, i) = Voronoi(i)%xyz(:) ; vor_h(i) = Voronoi(i)%h
          vor_first(i) = Voronoi(i)%first_neighbour ; vor_last(i) = Voronoi(i)%last_neighbour
          vor_cut(i) = l2i(Voronoi(i)%was_cut) ; vor_star(i) = l2i(Voronoi(i)%is_star) ; vor_star_nb(i) = l2i(Voronoi(i)%is_star_neighbour)
       enddo
       g%vor_xyz = c_loc(vor_xyz) ; g%vor_h = c_loc(vor_h) ; g%vor_first = c_loc(vor_first) ; g%vor_last = c_loc(vor_last)
       g%vor_was_cut = c_loc(vor_cut) ; g%vor_is_star = c_loc(vor_star) ; g%vor_is_star_neighbour = c_loc(vor_star_nb)
       g%neighbours_list = c_loc(neighbours_list) ; g%n_neighbours_tot = size(neighbours_list, kind=c_int64_t)
       do i = 1, 6
          g%wall_x(:, i) = (/ wall(i)%x1, wall(i)%x2, wall(i)%x3, wall(i)%x4 /)
       enddo
       g%cutting_distance_o_h = PS%cutting_distance_o_h
    endif
    g%volume = c_loc(volume)
    g%n_stars = n_stars
    g%star_xyzr = c_loc(star_xyzr) ; g%star_icell = c_loc(star_icell) ; g%star_out_model = c_loc(star_out)
    call b200_check(mcfost_b200_multi_upload_grid(b200, g), "upload_grid")

    o%n_lambda = n_lambda ; o%p_n_cells = p_n_cells ; o%p_n_lambda_pos = p_n_lambda_pos ; o%n_T = n_T
    o%kappa = c_loc(kappa) ; o%kappa_abs_LTE = c_loc(kappa_abs_LTE) ; o%kappa_factor = c_loc(kappa_factor)
    o%tab_albedo_pos = c_loc(tab_albedo_pos) ; o%tab_g_pos = c_loc(tab_g_pos)
    o%prob_s11_pos = c_null_ptr ; o%tab_s11_pos = c_null_ptr
    o%tab_s12_o_s11_pos = c_null_ptr ; o%tab_s22_o_s11_pos = c_null_ptr ; o%tab_s33_o_s11_pos = c_null_ptr
    o%tab_s34_o_s11_pos = c_null_ptr ; o%tab_s44_o_s11_pos = c_null_ptr
    if (allocated(prob_s11_pos)) o%prob_s11_pos = c_loc(prob_s11_pos)      ! scattering method 2 only (mem.f90:223)
    if (allocated(tab_s11_pos)) o%tab_s11_pos = c_loc(tab_s11_pos)
    if (lsepar_pola .and. allocated(tab_s12_o_s11_pos)) then
       o%tab_s12_o_s11_pos = c_loc(tab_s12_o_s11_pos) ; o%tab_s22_o_s11_pos = c_loc(tab_s22_o_s11_pos)
       o%tab_s33_o_s11_pos = c_loc(tab_s33_o_s11_pos) ; o%tab_s34_o_s11_pos = c_loc(tab_s34_o_s11_pos)
       o%tab_s44_o_s11_pos = c_loc(tab_s44_o_s11_pos)
    endif
    o%log_Qcool_minus_extra_heating = c_loc(log_Qcool_minus_extra_heating) ; o%kdB_dT_CDF = c_loc(kdB_dT_CDF)
    o%tab_Temp = c_loc(tab_Temp) ; o%T_min = T_min
    call b200_check(mcfost_b200_multi_upload_opacity(b200, o), "upload_opacity")
    call b200_upload_dark_zone()
  end subroutine b200_upload_model

  ! after define_dark_zone (per wavelength in the SED step, dust_transfer.f90:919)
  subroutine b200_upload_dark_zone()
    if (allocated(dark_i32)) deallocate(dark_i32)
    allocate(dark_i32(n_cells))
    dark_i32 = merge(1_c_int32_t, 0_c_int32_t, l_dark_zone(1:n_cells))
    call b200_check(mcfost_b200_multi_upload_dark_zone(b200, dark_i32), "upload_dark_zone")
  end subroutine b200_upload_dark_zone

  ! Replaces  call define_dark_zone(lambda, p_lambda, tau_max, ldiff_approx)  (optical_depth.f90:1425-1651; call sites
  ! dust_transfer.f90:293,919) on structured grids: GPU 0 of the node does the sums and the ray walks, the result goes
  ! into the reference's own module variables and is installed on every GPU.
  subroutine define_dark_zone_b200(lambda, p_lambda, tau_max, ldiff_approx)
    integer, intent(in) :: lambda, p_lambda
    real, intent(in) :: tau_max
    logical, intent(in) :: ldiff_approx
    integer(c_int32_t), allocatable :: iRmin(:), iRmax(:), zinf(:,:)
    real(c_double), allocatable, target :: dust_sum(:)
    integer(c_int32_t) :: flag
    type(c_ptr) :: pds
    integer :: i
    allocate(iRmin(max(n_regions,1)), iRmax(max(n_regions,1)))
    do i = 1, n_regions
       iRmin(i) = regions(i)%iRmin ; iRmax(i) = regions(i)%iRmax
    enddo
    pds = c_null_ptr
    if (n_zones > 1) then
       allocate(dust_sum(n_cells))
       do i = 1, n_cells
          dust_sum(i) = sum(dust_density_o_n_grains(:,i))
       enddo
       pds = c_loc(dust_sum)
    endif
    if (allocated(dark_i32)) deallocate(dark_i32)
    allocate(dark_i32(n_cells))
    if (l3D) then
       call b200_check(mcfost_b200_define_dark_zone(mcfost_b200_multi_handle(b200, 0_c_int), int(lambda, c_int32_t), tau_max, &
            r_grid, z_grid, int(n_regions, c_int32_t), iRmin, iRmax, pds, dark_i32, ri_in_dark_zone, ri_out_dark_zone, &
            zj_sup_dark_zone, zj_inf_dark_zone, flag), "define_dark_zone")
    else      ! zj_inf_dark_zone is not used on a 2D grid (a scratch array of the same shape is passed)
       allocate(zinf(n_rad, n_az)) ; zinf = 0
       call b200_check(mcfost_b200_define_dark_zone(mcfost_b200_multi_handle(b200, 0_c_int), int(lambda, c_int32_t), tau_max, &
            r_grid, z_grid, int(n_regions, c_int32_t), iRmin, iRmax, pds, dark_i32, ri_in_dark_zone, ri_out_dark_zone, &
            zj_sup_dark_zone, zinf, flag), "define_dark_zone")
    endif
    l_dark_zone(1:n_cells) = dark_i32 /= 0
    l_is_dark_zone = flag /= 0
    if ((ldiff_approx).and.(n_rad > 1)) then      ! optical_depth.f90:1629-1632
       if (minval(ri_in_dark_zone(:))==1) call error("first cell is in diffusion approximation zone", &
            msg2="Increase spatial grid resolution")
    endif
    call b200_check(mcfost_b200_multi_upload_dark_zone(b200, dark_i32), "upload_dark_zone")      ! the other GPUs of the node
  end subroutine define_dark_zone_b200

  ! Replaces  call repartition_energie(lambda)  (thermal_emission.f90:1771-1949) for the LTE case without lweight_emission:
  ! every GPU builds prob_E_cell(:,lambda), frac_E_stars(lambda), frac_E_disk(lambda) where the photon loop reads them (the
  ! host copy of prob_E_cell is only fetched from GPU 0 when the caller needs it: lwant_prob); E_disk, the fractions and
  ! E_totale come back as the reference computes them.  mc_photon_loop_b200 then passes NULL for the three tables.
  subroutine repartition_energie_b200(lambda, lwant_prob)
    integer, intent(in) :: lambda
    logical, intent(in) :: lwant_prob
    real(c_double) :: surface
    type(c_ptr) :: pp
    integer :: i
    if (.not.lRE_LTE .or. lRE_nLTE .or. lnRE .or. lweight_emission) &
         call error("mcfost_b200: repartition_energie_b200 covers the LTE case without lweight_emission")
    do i = 0, mcfost_b200_multi_n_gpus(b200) - 1
       pp = c_null_ptr
       ! (the C entry point takes the base of prob_E_cell(0:n_cells, n_lambda) and fills the columns of its wavelength range)
       if (i == 0 .and. lwant_prob) pp = c_loc(prob_E_cell(0,1))
       call b200_check(mcfost_b200_repartition_energie(mcfost_b200_multi_handle(b200, int(i, c_int)), int(lambda, c_int32_t), &
            int(lambda, c_int32_t), Tdust, tab_lambda, E_stars, E_ISM, c_null_ptr, E_disk, frac_E_stars, frac_E_disk, c_null_ptr, pp), &
            "repartition_energie")
    enddo
    surface=4*pi*(pc_to_AU*distance)**2
    if (l_sym_centrale) then
       E_totale(lambda) = 2.0*pi*hp*c_light**2/surface * (E_stars(lambda)+E_disk(lambda)+E_ISM(lambda)) * real(N_thet)*real(N_phi)
    else
       E_totale(lambda) = 2.0*pi*hp*c_light**2/surface * (E_stars(lambda)+E_disk(lambda)+E_ISM(lambda)) * real(2*N_thet)*real(N_phi)
    endif
  end subroutine repartition_energie_b200

  ! Replaces  call integ_tau(lambda)  (optical_depth.f90:186-244): the optical depth from the star through the midplane and
  ! along the inclination of interest, two rays through mcfost_b200_optical_length_tot, same messages.
  subroutine integ_tau_b200(lambda)
    integer, intent(in) :: lambda
    real(c_double) :: x0(2), y0(2), z0(2), u0(2), v0(2), w0(2), tau(2), lmin(2), lmax(2)
    integer(c_int32_t) :: icell(2)
    type(c_ptr) :: h0
    integer :: k, ic
    h0 = mcfost_b200_multi_handle(b200, 0_c_int)
    x0 = 0.0 ; y0 = 0.0 ; z0 = 0.0 ; v0 = 0.0
    u0(1) = 1.0 ; w0(1) = 0.0
    w0(2) = cos((angle_interet)*pi/180.) ; u0(2) = sqrt(1.0-w0(2)*w0(2))
    call b200_check(mcfost_b200_index_cell(h0, 2_c_int64_t, x0, y0, z0, icell), "index_cell")
    call b200_check(mcfost_b200_optical_length_tot(h0, 2_c_int64_t, int(lambda, c_int32_t), x0, y0, z0, u0, v0, w0, icell, &
         tau, lmin, lmax, c_null_ptr), "optical_length_tot")
    do k = 1, 2
       if (k == 1) then
          write(*,*) 'Integ tau in midplane = ', real(tau(1))
       else
          write(*,fmt='(" Integ tau (i =",f4.1," deg)   = ",E12.5)') angle_interet, real(tau(2))
       endif
       if (.not.lvariable_dust) then
          ic = icell_not_empty
          if (kappa(icell1,lambda) * kappa_factor(ic) > tiny_real) then
             write(*,*) " Column density (g/cm^2)   = ", real(tau(k)*(dust_mass(ic)/(volume(ic)*AU_to_cm**3))/ &
                  (kappa(icell1,lambda) * kappa_factor(ic)/AU_to_cm))
          endif
       endif
    enddo
  end subroutine integ_tau_b200

  ! Replaces  call compute_column(type, column, lambda)  (optical_depth.f90:328-415) on GPU 0: type 2 = optical depth at
  ! lambda, types 1 / 3 = (molecular) column density with the per-cell weight formed here exactly as the reference forms it.
  subroutine compute_column_b200(type, column, lambda)
    use density, only : gas_density
    use molecular_emission, only : tab_abundance
    integer, intent(in) :: type
    integer, intent(in), optional :: lambda
    real, dimension(n_cells,4), intent(out) :: column
    real(c_double), allocatable, target :: cx(:), cy(:), cz(:), factor(:)
    real(c_double) :: CD_units
    type(c_ptr) :: pf
    integer :: icell, lam
    allocate(cx(n_cells), cy(n_cells), cz(n_cells))
    do icell = 1, n_cells
       if (lVoronoi) then
          cx(icell) = Voronoi(icell)%xyz(1) ; cy(icell) = Voronoi(icell)%xyz(2) ; cz(icell) = Voronoi(icell)%xyz(3)
       else
          cx(icell) = r_grid(icell) * cos(phi_grid(icell)) ; cy(icell) = r_grid(icell) * sin(phi_grid(icell)) ; cz(icell) = z_grid(icell)
       endif
    enddo
    pf = c_null_ptr ; lam = 1
    if (type == 2) then
       lam = lambda
    else
       allocate(factor(n_cells))
       if (type == 1) then
          CD_units = AU_to_m * mu_mH / (m_to_cm)**2
          factor(:) = CD_units * gas_density(1:n_cells)
       else
          CD_units = AU_to_m / (m_to_cm)**2
          factor(:) = CD_units * gas_density(1:n_cells) * tab_abundance(1:n_cells)
       endif
       pf = c_loc(factor)
    endif
    call b200_check(mcfost_b200_compute_column(mcfost_b200_multi_handle(b200, 0_c_int), int(lam, c_int32_t), pf, cx, cy, cz, column), &
         "compute_column")
  end subroutine compute_column_b200

  ! Replaces the LTE part of  call init_reemission(lheating, dudt)  (thermal_emission.f90:404-550, high-memory branch, no
  ! extra heating) after b200_upload_model: every GPU builds log_Qcool_minus_extra_heating and kdB_dT_CDF from the
  ! kappa_abs_LTE it already holds; GPU 0's copies come back into the reference's module arrays (Temp_finale, outputs).
  ! With cell-dependent dust kdB_dT_CDF is n_lambda x n_T x n_cells doubles that no longer cross PCIe once per GPU.
  subroutine init_reemission_b200()
    integer :: i
    type(c_ptr) :: pq, pc
    if (lextra_heating .or. low_mem_th_emission) call error("mcfost_b200: init_reemission_b200 covers the high-memory LTE branch without extra heating")
    do i = 0, mcfost_b200_multi_n_gpus(b200) - 1
       pq = c_null_ptr ; pc = c_null_ptr
       if (i == 0) then
          pq = c_loc(log_Qcool_minus_extra_heating) ; pc = c_loc(kdB_dT_CDF)
       endif
       call b200_check(mcfost_b200_init_reemission(mcfost_b200_multi_handle(b200, int(i, c_int)), tab_lambda, tab_delta_lambda, pq, pc), &
            "init_reemission")
    enddo
  end subroutine init_reemission_b200

  ! after init_reemission / opacite when a per-grain mode is on, and again after every update_proba_abs_nRE
  ! (thermal_emission.f90:1518), which changes l_RE, kappa_abs_RE and the three probabilities
  subroutine b200_upload_grains()
    type(mcb_grains) :: q
    integer :: k
    q%n_grains_tot = n_grains_tot ; q%n_dens = size(dust_density_o_n_grains, 1)
    q%grain_RE_LTE_start = grain_RE_LTE_start ; q%grain_RE_LTE_end = grain_RE_LTE_end
    q%grain_RE_nLTE_start = grain_RE_nLTE_start ; q%grain_RE_nLTE_end = grain_RE_nLTE_end
    q%grain_nRE_start = grain_nRE_start ; q%grain_nRE_end = grain_nRE_end
    if (allocated(grain_zone_i32)) deallocate(grain_zone_i32)
    allocate(grain_zone_i32(n_grains_tot))
    do k = 1, n_grains_tot
       grain_zone_i32(k) = grain(k)%zone
    enddo
    q%grain_zone = c_loc(grain_zone_i32) ; q%n_grains = c_loc(nbre_grains) ; q%dust_density_o_n_grains = c_loc(dust_density_o_n_grains)
    q%C_abs = c_loc(C_abs) ; q%C_abs_norm = c_loc(C_abs_norm) ; q%C_sca = c_loc(C_sca) ; q%tab_g = c_loc(tab_g)
    q%prob_s11 = c_null_ptr ; q%tab_s11 = c_null_ptr ; q%tab_s12 = c_null_ptr ; q%tab_s22 = c_null_ptr
    q%tab_s33 = c_null_ptr ; q%tab_s34 = c_null_ptr ; q%tab_s44 = c_null_ptr ; q%ksca_CDF = c_null_ptr
    if (lscattering_method1) then
       q%prob_s11 = c_loc(prob_s11) ; q%tab_s11 = c_loc(tab_s11)
       if (lsepar_pola) then
          q%tab_s12 = c_loc(tab_s12) ; q%tab_s22 = c_loc(tab_s22) ; q%tab_s33 = c_loc(tab_s33) ; q%tab_s34 = c_loc(tab_s34) ; q%tab_s44 = c_loc(tab_s44)
       endif
       if (.not.low_mem_scattering) q%ksca_CDF = c_loc(ksca_CDF)
    endif
    q%kappa_abs_nLTE = c_null_ptr ; q%kabs_nLTE_CDF = c_null_ptr ; q%log_E_em_1grain = c_null_ptr ; q%kdB_dT_1grain_nLTE_CDF = c_null_ptr
    if (lRE_nLTE) then
       q%kappa_abs_nLTE = c_loc(kappa_abs_nLTE) ; q%log_E_em_1grain = c_loc(log_E_em_1grain) ; q%kdB_dT_1grain_nLTE_CDF = c_loc(kdB_dT_1grain_nLTE_CDF)
       if (.not.low_mem_th_emission_nLTE) q%kabs_nLTE_CDF = c_loc(kabs_nLTE_CDF)
    endif
    q%kappa_abs_RE = c_null_ptr ; q%proba_abs_RE = c_null_ptr ; q%Proba_abs_RE_LTE = c_null_ptr ; q%Proba_abs_RE_LTE_p_nLTE = c_null_ptr
    q%log_E_em_1grain_nRE = c_null_ptr ; q%kdB_dT_1grain_nRE_CDF = c_null_ptr ; q%l_RE = c_null_ptr ; q%J0 = c_null_ptr
    if (lnRE) then
       if (allocated(l_RE_i32)) deallocate(l_RE_i32)
       allocate(l_RE_i32(grain_nRE_start:grain_nRE_end, n_cells))
       l_RE_i32 = merge(1_c_int32_t, 0_c_int32_t, l_RE)
       q%kappa_abs_RE = c_loc(kappa_abs_RE) ; q%proba_abs_RE = c_loc(proba_abs_RE)
       q%log_E_em_1grain_nRE = c_loc(log_E_em_1grain_nRE) ; q%kdB_dT_1grain_nRE_CDF = c_loc(kdB_dT_1grain_nRE_CDF) ; q%l_RE = c_loc(l_RE_i32)
    endif
    if (.not.lonly_LTE) then
       q%Proba_abs_RE_LTE = c_loc(Proba_abs_RE_LTE) ; q%Proba_abs_RE_LTE_p_nLTE = c_loc(Proba_abs_RE_LTE_p_nLTE) ; q%J0 = c_loc(J0)
    endif
    q%kdB_dT_1grain_LTE_CDF = c_null_ptr
    if (low_mem_th_emission) q%kdB_dT_1grain_LTE_CDF = c_loc(kdB_dT_1grain_LTE_CDF)
    call b200_check(mcfost_b200_multi_upload_grains(b200, q), "upload_grains")
  end subroutine b200_upload_grains

  !------------------------------------------------------------------------------------------------------------------
  ! the drop-in: same dummy arguments as mc_photon_loop (dust_transfer.f90:439-454)
  subroutine mc_photon_loop_b200(lambda_in, p_lambda_in, n_photons2, n_phot_lim, nnfot1_start, laffichage)
    integer, intent(in) :: lambda_in, p_lambda_in, n_photons2, nnfot1_start
    real, intent(in) :: n_phot_lim
    logical, intent(in) :: laffichage
    type(mcb_emission) :: e ; type(mcb_run_params) :: r ; type(mcb_tallies) :: t

    ! emission tables change every temperature iteration / wavelength (repartition_energie, thermal_emission.f90:1771)
    e%spectre_emission_cumul = c_loc(spectre_emission_cumul) ; e%frac_E_stars = c_loc(frac_E_stars) ; e%frac_E_disk = c_loc(frac_E_disk)
    e%prob_E_cell = c_loc(prob_E_cell) ; e%CDF_E_star = c_loc(CDF_E_star)
    e%L_packet_th = L_packet_th ; e%E_paquet = E_paquet ; e%R_ISM = R_ISM ; e%centre_ISM = centre_ISM
    e%correct_E_emission = c_null_ptr
    if (lweight_emission) e%correct_E_emission = c_loc(correct_E_emission)
    call b200_check(mcfost_b200_multi_upload_emission(b200, e), "upload_emission")

    r%lambda_in = lambda_in ; r%p_lambda_in = p_lambda_in ; r%n_photons2 = n_photons2
    r%n_phot_lim = n_phot_lim
    r%nnfot1_start = nnfot1_start ; r%laffichage = l2i(laffichage)
    r%n_photons_loop = n_photons_loop
    r%letape_th = l2i(letape_th) ; r%lmono = l2i(lmono) ; r%lmono0 = l2i(lmono0)
    r%lscatt_ray_tracing1 = l2i(lscatt_ray_tracing1) ; r%lscatt_ray_tracing2 = l2i(lscatt_ray_tracing2)
    r%lsepar_pola = l2i(lsepar_pola) ; r%lsepar_contrib = l2i(lsepar_contrib)
    r%lscattering_method1 = l2i(lscattering_method1) ; r%lmethod_aniso1 = l2i(lmethod_aniso1) ; r%lisotropic = l2i(lisotropic)
    r%l_sym_centrale = l2i(l_sym_centrale) ; r%l_sym_axiale = l2i(l_sym_axiale)
    r%lonly_LTE = l2i(lonly_LTE) ; r%lxJ_abs_step1 = l2i(lxJ_abs_step1) ; r%lxJ_abs = l2i(lxJ_abs)
    r%N_thet = N_thet ; r%N_phi = N_phi ; r%capt_sup = capt_sup
    r%RT_n_incl = RT_n_incl ; r%RT_n_az = RT_n_az
    r%tab_u_rt = c_null_ptr ; r%tab_v_rt = c_null_ptr ; r%tab_w_rt = c_null_ptr
    if (lscatt_ray_tracing1 .and. .not.letape_th) then
       r%tab_u_rt = c_loc(tab_u_rt) ; r%tab_v_rt = c_loc(tab_v_rt) ; r%tab_w_rt = c_loc(tab_w_rt)
    endif
    r%seed = int(seed, c_int64_t)
    r%call_index = b200_call_index ; b200_call_index = b200_call_index + 1
    r%rank = 0 ; r%n_ranks = 1                  ! set by mcfost_b200_multi_run (GPU g = rank g)
    ! tallies accumulate over the calls of one step exactly where the reference accumulates them: the thermal step is one
    ! call; in the SED / image steps every wavelength writes its own (lambda, ...) slots
    r%reset_tallies = l2i(letape_th .or. lambda_in == 1)
    r%loutput_mc = l2i(loutput_mc) ; r%n_theta_I = n_theta_I ; r%n_phi_I = n_phi_I
    r%lonly_nLTE = l2i(lonly_nLTE) ; r%lRE_nLTE = l2i(lRE_nLTE) ; r%lnRE = l2i(lnRE)
    r%low_mem_th_emission_nLTE = l2i(low_mem_th_emission_nLTE) ; r%low_mem_scattering = l2i(low_mem_scattering)
    r%npix_x = npix_x ; r%npix_y = npix_y
    r%zoom = zoom
    r%map_size = map_size ; r%cos_disk = cos_disk ; r%sin_disk = sin_disk
    r%l_sym_ima = l2i(l_sym_ima) ; r%lonly_capt_interet = l2i(lonly_capt_interet) ; r%capt_inf = capt_inf
    r%lorigine = l2i(lorigine) ; r%capt_interet = capt_interet
    r%low_mem_th_emission = l2i(low_mem_th_emission) ; r%lweight_emission = l2i(lweight_emission) ; r%lspot = l2i(lspot)
    r%T_spot = T_spot ; r%surf_fraction_spot = surf_fraction_spot ; r%theta_spot = theta_spot ; r%phi_spot = phi_spot
    r%star1_T = star(1)%T
    r%tab_lambda = c_loc(tab_lambda)
    r%lxN_abs = l2i((letape_th .and. lmcfost_lib) .or. (.not.letape_th .and. lProDiMo))      ! radiation_field.f90:53,60
    r%lMRW = l2i(lMRW) ; r%gamma_MRW = 2.0                                                   ! MRW.f90:11
    r%lcount_sent = l2i((lProDiMo .or. lML) .and. .not.letape_th)                            ! dust_transfer.f90:512-516
    r%max_inflight_fraction = 0.0
    r%lISM_loop = 0            ! 1 in the call that replaces the ISM side loop of run_sed_mc (dust_transfer.f90:941-985)

    ! the id = 1 slices receive the merged tallies; the other slices are zeroed so that the untouched Fortran reducers
    ! (sum(xKJ_abs(icell,:)) thermal_emission.f90:668, sum(sed(lambda,:,:,:),dim=3) output.f90:3102, sum(n_phot_envoyes(lambda,:))
    ! :3084, sum(xI_scatt(...,:)) dust_ray_tracing.f90:689) keep working
    t%xKJ_abs = c_null_ptr ; t%xJ_abs = c_null_ptr ; t%xI_scatt = c_null_ptr ; t%I_spec = c_null_ptr ; t%I_spec_star = c_null_ptr
    t%xT_ech_1grain = c_null_ptr ; t%xT_ech_1grain_nRE = c_null_ptr ; t%stokes_map = c_null_ptr
    t%star_origin = c_null_ptr ; t%disk_origin = c_null_ptr ; t%xN_abs = c_null_ptr
    if (allocated(xKJ_abs)) then
       if (nb_proc > 1) xKJ_abs(:, 2:) = 0.0_dp
       t%xKJ_abs = c_loc(xKJ_abs(1, 1))
    endif
    if (allocated(xJ_abs) .and. (r%lxJ_abs_step1 /= 0 .or. r%lxJ_abs /= 0)) then
       if (nb_proc > 1) xJ_abs(:, :, 2:) = 0.0_dp
       t%xJ_abs = c_loc(xJ_abs(1, 1, 1))
    endif
    t%xT_ech = c_loc(xT_ech(1, 1))
    if (nb_proc > 1) n_phot_envoyes(:, 2:) = 0.0_dp
    t%n_phot_envoyes = c_loc(n_phot_envoyes(1, 1))
    if (nb_proc > 1) then
       sed(:,:,:,2:) = 0.0_dp ; sed_q(:,:,:,2:) = 0.0_dp ; sed_u(:,:,:,2:) = 0.0_dp ; sed_v(:,:,:,2:) = 0.0_dp ; n_phot_sed(:,:,:,2:) = 0.0_dp
       sed_star(:,:,:,2:) = 0.0_dp ; sed_star_scat(:,:,:,2:) = 0.0_dp ; sed_disk(:,:,:,2:) = 0.0_dp ; sed_disk_scat(:,:,:,2:) = 0.0_dp
    endif
    t%sed = c_loc(sed(1,1,1,1)) ; t%sed_q = c_loc(sed_q(1,1,1,1)) ; t%sed_u = c_loc(sed_u(1,1,1,1)) ; t%sed_v = c_loc(sed_v(1,1,1,1))
    t%n_phot_sed = c_loc(n_phot_sed(1,1,1,1))
    t%sed_star = c_loc(sed_star(1,1,1,1)) ; t%sed_star_scat = c_loc(sed_star_scat(1,1,1,1))
    t%sed_disk = c_loc(sed_disk(1,1,1,1)) ; t%sed_disk_scat = c_loc(sed_disk_scat(1,1,1,1))
    if (lscatt_ray_tracing1 .and. .not.letape_th) then
       if (nb_proc > 1) xI_scatt(:,:,:,:,:,2:) = 0.0
       t%xI_scatt = c_loc(xI_scatt(1,1,1,1,1,1))
    endif
    t%N_type_flux = N_type_flux
    if (lscatt_ray_tracing2 .and. .not.letape_th) then
       if (nb_proc > 1) then
          I_spec(:,:,:,:,2:) = 0.0 ; I_spec_star(:,2:) = 0.0
       endif
       t%I_spec = c_loc(I_spec(1,1,1,1,1)) ; t%I_spec_star = c_loc(I_spec_star(1,1))
    endif
    t%stats = c_loc(stats)
    if (lRE_nLTE) t%xT_ech_1grain = c_loc(xT_ech_1grain(grain_RE_nLTE_start, 1, 1))
    if (lnRE) t%xT_ech_1grain_nRE = c_loc(xT_ech_1grain_nRE(grain_nRE_start, 1, 1))
    t%E_abs_nRE = c_loc(E_abs_nRE_c)
    if (lmono0 .and. loutput_mc) t%stokes_map = c_null_ptr      ! the maps of ONE wavelength: copied into STOKEI(lambda_in,...) etc. by the caller, see INTEGRATION.md
    if (lorigine) then
       t%star_origin = c_loc(star_origin(1, 1)) ; t%disk_origin = c_loc(disk_origin(1, 1, 1))
    endif

    call b200_check(mcfost_b200_multi_run(b200, r, t), "run")

    E_abs_nRE = E_abs_nRE_c
    ! Temp_LTE(id=0) takes minval(xT_ech(icell,:)) (thermal_emission.f90:683); the per-grain readers take maxval(...,:) (:823,:977)
    if (nb_proc > 1) then
       xT_ech(:, 2:) = spread(xT_ech(:, 1), 2, nb_proc - 1)
       if (lRE_nLTE) xT_ech_1grain(:, :, 2:) = spread(xT_ech_1grain(:, :, 1), 3, nb_proc - 1)
       if (lnRE) xT_ech_1grain_nRE(:, :, 2:) = spread(xT_ech_1grain_nRE(:, :, 1), 3, nb_proc - 1)
    endif
  end subroutine mc_photon_loop_b200

  subroutine b200_finalize()
    if (c_associated(b200)) call mcfost_b200_multi_finalize(b200)
    b200 = c_null_ptr
  end subroutine b200_finalize

end module mcfost_b200_shim
