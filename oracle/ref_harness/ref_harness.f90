! ref_harness.f90 -- lets a maintainer PIN the oracle to the reference itself.
!
! Links the reference's own grid routines (cylindrical_grid.f90 and what it uses) and dumps, for seeded rays,
!   cross_cylindrical_cell (cylindrical_grid.f90:918-1175), index_cell (:833-890), move_to_grid_cyl (:1284-1411),
!   distance_to_closest_wall_cyl (:1179-1226)
! to a stream file that oracle/ref_harness/to_npz.py turns into tests/golden/ref_vectors_cyl.npz; tests/test_ref_vectors.py
! then requires the oracle (and, with -m gpu, the CUDA kernels) to reproduce every vector: bit-exact cell indices, lengths
! to 1e-12.  No Fortran compiler exists in the environment this file was written in: it is provided ready to build
! (see Makefile), unbuilt and untested.
!
! Input (written by make_inputs.py): a stream file with
!   int32 n_rad, nz, n_az, l3D, n_rays ; real64 r_lim(0:n_rad), r_lim_2(0:n_rad), r_lim_3(0:n_rad), z_lim(n_rad,nz+2),
!   zmax(n_rad), tan_phi_lim(n_az), sin_phi_lim(n_az), cos_phi_lim(n_az), Rmax2, zmaxmax ;
!   real64 x(n), y(n), z(n), u(n), v(n), w(n) ; int32 icell(n)
program ref_harness
  use mcfost_env, only : dp
  use parameters
  use constants
  use messages
  use cylindrical_grid
  implicit none

  integer :: nrays, i, l3D_i, next_cell, icell_in, lint
  real(kind=dp), allocatable :: x(:), y(:), z(:), u(:), v(:), w(:)
  integer, allocatable :: icell(:)
  real(kind=dp) :: x1, y1, z1, l, l_contrib, l_void, xx, yy, zz
  logical :: lintersect
  character(len=512) :: fin, fout

  call get_command_argument(1, fin) ; call get_command_argument(2, fout)
  open(unit=10, file=trim(fin), access='stream', form='unformatted', status='old')
  read(10) n_rad, nz, n_az, l3D_i, nrays
  l3D = (l3D_i /= 0) ; lcylindrical = .true. ; lspherical = .false. ; lVoronoi = .false.
  n_zones = 1 ; n_rad_in = 1
  if (l3D) then
     j_start = -nz ; n_cells = 2 * n_rad * nz * n_az
  else
     j_start = 1 ; n_cells = n_rad * nz
  endif
  allocate(r_lim(0:n_rad), r_lim_2(0:n_rad), r_lim_3(0:n_rad), z_lim(n_rad, nz + 2), zmax(n_rad))
  allocate(tan_phi_lim(n_az), sin_phi_lim(n_az), cos_phi_lim(n_az))
  read(10) r_lim, r_lim_2, r_lim_3, z_lim, zmax, tan_phi_lim, sin_phi_lim, cos_phi_lim, Rmax2, zmaxmax
  allocate(x(nrays), y(nrays), z(nrays), u(nrays), v(nrays), w(nrays), icell(nrays))
  read(10) x, y, z, u, v, w, icell
  close(10)

  call build_cylindrical_cell_mapping()            ! cylindrical_grid.f90:45-179: cell_map, cell_map_i/j/k, lexit_cell

  open(unit=11, file=trim(fout), access='stream', form='unformatted', status='replace')
  write(11) nrays
  do i = 1, nrays
     ! cross_cylindrical_cell(x0,y0,z0, u,v,w, cell, previous_cell, x1,y1,z1, next_cell, l, l_contrib, l_void_before)
     call cross_cylindrical_cell(x(i), y(i), z(i), u(i), v(i), w(i), icell(i), 0, x1, y1, z1, next_cell, l, l_contrib, l_void)
     write(11) x1, y1, z1, l, next_cell
     call index_cell_cyl(x(i), y(i), z(i), icell_in)                        ! cylindrical_grid.f90:833
     write(11) icell_in
     write(11) distance_to_closest_wall_cyl(icell(i), x(i), y(i), z(i))
     ! entry from outside: start 3 Rmax away, flying towards the ray's own point
     xx = x(i) - 3.0_dp * sqrt(Rmax2) * u(i) ; yy = y(i) - 3.0_dp * sqrt(Rmax2) * v(i) ; zz = z(i) - 3.0_dp * sqrt(Rmax2) * w(i)
     call move_to_grid_cyl(0, xx, yy, zz, u(i), v(i), w(i), icell_in, lintersect)
     lint = merge(1, 0, lintersect)
     write(11) xx, yy, zz, icell_in, lint
  enddo
  close(11)
end program ref_harness
