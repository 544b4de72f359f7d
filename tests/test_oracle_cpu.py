"""CPU tests of the oracle (test infrastructure) -- run with -m "not gpu".

The reference holds no golden vectors for these routines, so the
oracle is pinned by (a) analytic properties, (b) Random123's published Philox
known-answer vectors and (c) its own committed golden vectors (regression)."""
import os

import numpy as np
import pytest

from mcfost_b200 import synthetic as S
from oracle import binding
from oracle.binding import Oracle

from helpers import rays_in_cells, rays_from_outside, small_problems

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert list(binding.philox([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(binding.philox([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(binding.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_rng_stream_uniform():
    u = binding.rng_stream(269753, 0, 12345, 200000)
    assert (u >= 0).all() and (u < 1).all()
    assert abs(u.mean() - 0.5) < 3e-3 and abs(u.var() - 1 / 12) < 2e-3
    # streams of different packets / calls differ
    assert not np.array_equal(u[:16], binding.rng_stream(269753, 0, 12346, 16))
    assert not np.array_equal(u[:16], binding.rng_stream(269753, 1, 12345, 16))


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D", "sph3D"])
def test_cell_numbering_matches_reference_builder(name):
    """build_cylindrical_cell_mapping restated twice (numpy generator and C++ oracle)."""
    P = small_problems()[name]()
    O = Oracle(P)          # set_grid cross-checks the maps and fails with MCB_ERR_CELL_MAP otherwise
    ci, cj, ck, lexit = O.cell_maps()
    assert np.array_equal(ci, P.cell_map_i) and np.array_equal(cj, P.cell_map_j) and np.array_equal(ck, P.cell_map_k)
    assert O.n_cells_tot == P.n_cells_tot
    # exit flags: radial virtual shell = 1, top/bottom rows = 2 (cylindrical_grid.f90:131-132)
    assert (lexit[ci == P.n_rad + 1] == 1).all()
    assert (lexit[:P.n_cells] == 0).all()
    assert O.index_cell([0.0], [0.0], [0.0])[0] == P.star_icell[0]


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D", "sph3D"])
def test_index_cell_round_trip(name):
    P = small_problems()[name]()
    O = Oracle(P)
    ic, x, y, z, *_ = rays_in_cells(P, 20000, seed=1)
    assert np.array_equal(O.index_cell(x, y, z), ic)


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D", "sph3D"])
def test_cross_cell_lands_on_a_wall_of_the_cell(name):
    """Geometric property: the exit point of a crossing lies on the cell boundary and
    index_cell of a point slightly beyond it is the announced next cell."""
    P = small_problems()[name]()
    O = Oracle(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 20000, seed=2)
    c = O.cross_cell(x, y, z, u, v, w, ic)
    assert (c["l"] > 0).all() and np.array_equal(c["l"], c["l_contrib"]) and (c["l_void_before"] == 0).all()
    ci, cj = P.cell_map_i[ic - 1], np.abs(P.cell_map_j[ic - 1])
    x1, y1, z1 = c["x1"], c["y1"], c["z1"]
    if P.kind == 1:
        r2 = x1 * x1 + y1 * y1
        d_r = np.minimum(np.abs(r2 / P.r_lim_2[ci] - 1), np.abs(r2 / P.r_lim_2[ci - 1] - 1))
        d_z = np.minimum(np.abs(np.abs(z1) - P.z_lim[ci - 1, cj]), np.abs(np.abs(z1) - P.z_lim[ci - 1, cj - 1])) / P.zmax[ci - 1]
        on_wall = (d_r < 1e-9) | (d_z < 1e-9)
    else:
        r2 = x1 * x1 + y1 * y1 + z1 * z1
        d_r = np.minimum(np.abs(r2 / P.r_lim_2[ci] - 1), np.abs(r2 / P.r_lim_2[ci - 1] - 1))
        tt = np.abs(z1) / np.sqrt(x1 * x1 + y1 * y1)
        d_t = np.minimum(np.abs(tt / P.tan_theta_lim[cj] - 1), np.abs(tt / np.maximum(P.tan_theta_lim[cj - 1], 1e-300) - 1))
        on_wall = (d_r < 1e-5) | (d_t < 1e-5) | (tt < 1e-9)      # tan_theta_lim(0) = 1e-10: the equatorial "cone"
    if P.l3D:
        phi = np.mod(np.arctan2(y1, x1), 2 * np.pi) / (2 * np.pi) * P.n_az
        on_wall |= np.abs(phi - np.round(phi)) < 1e-5
    assert on_wall.mean() > 0.9999
    # the announced next cell is where a point just past the wall is located
    eps = 1e-4 * np.sqrt(x1 * x1 + y1 * y1 + z1 * z1)   # > the fp32 resolution of the z index (cylindrical_grid.f90:868)
    nxt = O.index_cell(x1 + eps * u, y1 + eps * v, z1 + eps * w)
    real = c["next_cell"] <= P.n_cells
    assert (nxt[real] == c["next_cell"][real]).mean() > 0.98


def test_optical_length_tot_matches_analytic_column():
    """A radial midplane-parallel ray from inside the inner edge: tau = sum kappa*kappa_factor*dr."""
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    z0 = 1e-3 * P.zmax[0]
    r = O.optical_length_tot(P.lambda_seuil, [0.5 * P.r_lim[0]], [0.0], [z0], [1.0], [0.0], [0.0], O.index_cell([0.5 * P.r_lim[0]], [0.0], [z0]))
    mid = np.arange(P.n_rad)          # cells (i, 1)
    tau = float(np.sum(P.kappa[0, P.lambda_seuil - 1] * P.kappa_factor[mid] * (P.r_lim[1:] - P.r_lim[:-1])))
    assert abs(r["tau_tot"][0] - tau) / tau < 1e-6
    assert abs(r["lmax"][0] - (P.r_lim[-1] - 0.5 * P.r_lim[0])) / P.r_lim[-1] < 1e-9
    assert r["n_steps"][0] == P.n_rad + 1


def test_path_lengths_sum_to_chord():
    """Sum of per-cell crossing lengths along a ray = geometric chord to the outer cylinder."""
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 2000, seed=3)
    r = O.optical_length_tot(1, x, y, z, u, v, w, ic)
    # chord to cylinder R=Rout or the slab |z| = zmaxmax... reflect: the 2D grid mirrors z, so only the cylinder and
    # the top plane bound the walk
    a = u * u + v * v
    b = (x * u + y * v) / a
    cc = (x * x + y * y - P.Rmax2) / a
    s_cyl = -b + np.sqrt(b * b - cc)
    assert (r["lmax"] <= s_cyl * (1 + 1e-9)).all()
    assert (r["lmax"] > 0).all()


@pytest.mark.parametrize("name", ["cyl2D", "sph2D"])
def test_move_to_grid_enters_on_the_outer_boundary(name):
    P = small_problems()[name]()
    O = Oracle(P)
    x, y, z, u, v, w = rays_from_outside(P, 5000)
    m = O.move_to_grid(x, y, z, u, v, w)
    hit = m["lintersect"] == 1
    assert hit.mean() > 0.5
    r2 = m["x"] ** 2 + m["y"] ** 2 + (m["z"] ** 2 if P.kind == 2 else 0)
    on_cyl = np.abs(r2 / P.Rmax2 - 1) < 1e-8
    on_top = np.abs(np.abs(m["z"]) / max(P.zmaxmax, 1e-300) - 1) < 1e-8 if P.kind == 1 else np.zeros_like(on_cyl)
    assert (on_cyl | on_top)[hit].all()
    assert (m["icell"][hit] >= 1).all()


def test_samplers_unit_vectors_and_hg_mean():
    import ctypes as C
    lib = binding._load("liboracle.so")
    rng = np.random.default_rng(0)
    out = (C.c_double * 3)()
    for _ in range(200):
        w0 = rng.uniform(-1, 1); p0 = rng.uniform(0, 2 * np.pi)
        u0, v0 = np.sqrt(1 - w0 * w0) * np.cos(p0), np.sqrt(1 - w0 * w0) * np.sin(p0)
        cp = rng.uniform(-1, 1); ph = rng.uniform(-np.pi, np.pi)
        lib.oracle_cdapres(C.c_double(cp), C.c_double(ph), C.c_double(u0), C.c_double(v0), C.c_double(w0), out)
        assert abs(out[0] ** 2 + out[1] ** 2 + out[2] ** 2 - 1) < 1e-12
        assert abs(out[0] * u0 + out[1] * v0 + out[2] * w0 - cp) < 1e-9       # scattering angle preserved
    g = 0.6
    it, cs = C.c_int32(), C.c_double()
    mu = []
    for r in rng.uniform(0, 1, 20000).astype(np.float32):
        lib.oracle_hg(C.c_float(g), C.c_float(r), C.byref(it), C.byref(cs))
        mu.append(cs.value)
        assert 1 <= it.value <= 180
    assert abs(np.mean(mu) - g) < 0.01          # <cos theta> = g for Henyey-Greenstein


def test_thermal_energy_conservation_and_determinism():
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    a = O.run(n_threads=2, n_photons2=50)
    # every packet is either detected or killed (star hit); no energy created
    assert a.stats[0] == 128 * 50 == a.n_phot_envoyes.sum()
    assert a.sed.sum() == pytest.approx(a.stats[6])
    assert a.stats[5] + a.stats[6] == a.stats[0]
    assert a.n_phot_sed.sum() == a.stats[6]
    assert (a.sed_star + a.sed_star_scat + a.sed_disk + a.sed_disk_scat).sum() == pytest.approx(a.sed.sum())
    # per-packet Philox streams: the escaping SED is independent of the thread count up to the
    # running-temperature feedback; with one thread the run is exactly reproducible
    b = Oracle(P).run(n_threads=1, n_photons2=20)
    c = Oracle(P).run(n_threads=1, n_photons2=20)
    assert np.array_equal(b.xKJ_abs, c.xKJ_abs) and np.array_equal(b.sed, c.sed)


def test_optically_thin_grey_temperature_is_analytic():
    """Physics pin: grey, non-scattering, optically thin dust around a blackbody star reaches
    T(r) = T* sqrt(R*/(2r)) (dilute radiative equilibrium, e.g. Bjorkman & Wood 2001 eq. 1-3)."""
    P = S.spherical_shell(n_photons_eq_th=2000, tau_mid=1e-3, n_rad=20, nz=4)
    nl = P.n_lambda
    # make the dust grey and purely absorbing
    k0 = P.kappa[0, P.lambda_seuil - 1]
    P.kappa = np.asfortranarray(np.full((1, nl), k0)); P.kappa_abs_LTE = P.kappa.copy()
    P.tab_albedo_pos = np.asfortranarray(np.zeros((1, nl), np.float32))
    S.init_reemission(P); S.repartition_energie(P)
    t = Oracle(P, fast=True).run(n_threads=0, n_photons2=2000)
    T = S.temp_finale(P, t.xKJ_abs)
    r = np.sqrt(P.r_grid ** 2 + P.z_grid ** 2)
    T_an = 5000.0 * np.sqrt(P.star_xyzr[3, 0] / (2 * r))
    # the wavelength grid (50 bins) and temperature grid (100 bins) discretise B_nu: allow 3 %
    rel = np.abs(T / T_an - 1)
    assert np.median(rel) < 0.03, np.median(rel)


def test_sed_mode_forced_scattering_deterministic():
    """lmono: no temperature feedback, so thread count must not change any tally."""
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, lambda_in=5, p_lambda_in=5, n_photons2=10 ** 9, n_phot_lim=40.0,
              lscatt_ray_tracing1=1, RT_n_incl=2, RT_n_az=1,
              tab_u_rt=np.array([[0.5], [0.9]]), tab_v_rt=np.zeros((2, 1)), tab_w_rt=np.array([np.sqrt(0.75), np.sqrt(0.19)]))
    n_xI = 45 * 2 * 1 * 2 * P.n_cells
    a = Oracle(P).run(n_threads=1, n_xI=n_xI, **kw)
    b = Oracle(P).run(n_threads=4, n_xI=n_xI, **kw)
    assert a.stats[0] == 128 * 40
    assert np.allclose(a.sed, b.sed, rtol=1e-12, atol=0) and np.array_equal(a.n_phot_sed, b.n_phot_sed)
    assert np.allclose(a.xI_scatt, b.xI_scatt, rtol=1e-4, atol=1e-12)      # fp32 per-thread partial sums
    assert a.xI_scatt.sum() > 0
    # forced scattering: detected energy < sent packets, and strictly positive
    assert 0 < a.sed.sum() < a.stats[0]


def test_golden_vectors_regression():
    """The oracle's own golden vectors (generated by tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "oracle_golden.npz")
    assert os.path.exists(path), "run tests/golden/make_golden.py"
    G = np.load(path)
    for name in ("cyl2D", "cyl3D", "sph2D", "sph3D"):
        P = small_problems()[name]()
        O = Oracle(P)
        ic, x, y, z, u, v, w = rays_in_cells(P, 512, seed=11)
        c = O.cross_cell(x, y, z, u, v, w, ic)
        assert np.array_equal(c["next_cell"], G[f"{name}_next_cell"])
        assert np.array_equal(c["l"], G[f"{name}_l"])
        r = O.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
        assert np.array_equal(r["n_steps"], G[f"{name}_n_steps"])
        assert np.allclose(r["tau_tot"], G[f"{name}_tau"], rtol=1e-12, atol=0)
    P = small_problems()["cyl2D"]()
    t = Oracle(P).run(n_threads=1, n_photons2=5)
    assert np.array_equal(t.stats[:8], G["thermal_stats"])
    assert np.allclose(t.xKJ_abs, G["thermal_xKJ"], rtol=1e-12, atol=0)
    assert np.array_equal(t.sed, G["thermal_sed"])
    # per-grain branches and the complete capteur
    PG = S.multi_grain_like(n_photons_eq_th=5, n_rad=10, nz=6, n_rad_in=2, tau_mid=10.0)
    OG = Oracle(PG)
    t = OG.run(n_threads=1, xJ=True, n_photons2=5, lonly_LTE=0, lRE_nLTE=1, lnRE=1, lxJ_abs_step1=1)
    assert np.array_equal(t.stats[:8], G["mixed_stats"])
    assert np.allclose(t.xKJ_abs, G["mixed_xKJ"], rtol=1e-12, atol=0)
    assert np.array_equal(t.xT_ech_1grain, G["mixed_xT_1grain"]) and np.array_equal(t.xT_ech_1grain_nRE, G["mixed_xT_1grain_nRE"])
    assert np.allclose(t.E_abs_nRE, G["mixed_E_abs_nRE"], rtol=1e-12)
    t = OG.run(n_threads=1, letape_th=0, lmono=1, lambda_in=6, p_lambda_in=6, n_photons2=10 ** 9, n_phot_lim=5.0,
               lscattering_method1=1, lsepar_pola=1)
    assert np.array_equal(t.stats[:8], G["method1_stats"])
    assert np.allclose(t.sed, G["method1_sed"], rtol=1e-12, atol=0) and np.allclose(t.sed_q, G["method1_sed_q"], rtol=1e-10, atol=1e-300)
    t = OG.run(n_threads=1, letape_th=0, lmono=1, lmono0=1, loutput_mc=1, lambda_in=6, p_lambda_in=6, n_photons2=5,
               npix_x=8, npix_y=8, map_size=300.0, N_thet=3, N_phi=1, lsepar_pola=1, lsepar_contrib=1, l_sym_ima=1)
    assert np.allclose(t.stokes_map, G["maps_stokes"], rtol=1e-10, atol=1e-300)


def test_voronoi_mesh_and_walk():
    """Voronoi mesh stand-in (scipy + mirror trick): cells tile the box exactly; the oracle's
    fp32 plane walk (Voronoi.f90:839-992) goes from seed to wall through face-sharing cells."""
    P = S.voronoi_disk(n_points=800, n_photons_eq_th=20)
    assert abs(P.volume.sum() / 200.0 ** 3 - 1) < 1e-9
    O = Oracle(P)
    rng = np.random.default_rng(0)
    n = 3000
    ic = rng.integers(1, P.n_cells + 1, n).astype(np.int32)
    x, y, z = P.vor_xyz[0, ic - 1], P.vor_xyz[1, ic - 1], P.vor_xyz[2, ic - 1]
    assert np.array_equal(O.index_cell(x, y, z), ic)                       # nearest seed of a seed is itself
    w = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2 * np.pi, n)
    u, v = np.sqrt(1 - w * w) * np.cos(ph), np.sqrt(1 - w * w) * np.sin(ph)
    c = O.cross_cell(x, y, z, u, v, w, ic)
    assert (c["l"] > 0).all() and (c["next_cell"] != 0).all()
    # the next cell is a listed neighbour of the current one
    for k in range(200):
        nb = P.neighbours_list[P.vor_first[ic[k] - 1] - 1:P.vor_last[ic[k] - 1]]
        assert c["next_cell"][k] in nb
    # the exit point is equidistant (to fp32 accuracy) from the two seeds
    inner = c["next_cell"] > 0
    p1 = np.stack([c["x1"], c["y1"], c["z1"]])[:, inner]
    d0 = np.linalg.norm(p1 - P.vor_xyz[:, ic[inner] - 1], axis=0)
    d1 = np.linalg.norm(p1 - P.vor_xyz[:, c["next_cell"][inner] - 1], axis=0)
    assert np.median(np.abs(d0 - d1) / d0) < 1e-4
    # cut cells: contribution + void never exceed the crossing length
    assert (c["l_contrib"] <= c["l"] * (1 + 1e-12)).all() and (c["l_void_before"] <= c["l"] * (1 + 1e-12)).all()
    assert (c["l_contrib"][P.vor_was_cut[ic - 1] == 0] == c["l"][P.vor_was_cut[ic - 1] == 0]).all()
    # whole walks end on a wall: total length <= box diagonal
    r = O.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    assert (r["lmax"] > 0).all() and (r["lmax"] < 2 * np.sqrt(3) * 100.0).all()
    t = O.run(n_threads=2, n_photons2=20)
    assert t.stats[5] + t.stats[6] == t.stats[0] == 128 * 20
    assert t.sed.sum() == pytest.approx(t.stats[6])


def test_interstellar_radiation_field_packets():
    """emit_packet_ISM (stars.f90:728-787): packets start on a sphere of radius R_ISM around the model, fly inwards
    with a cosine law, enter through move_to_grid, heat the dust, and are never detected (flag_ISM)."""
    P = small_problems()["cyl2D"]()
    P.E_ISM = 0.5 * P.E_stars                      # a third of the energy comes from outside
    P.R_ISM = 1.5 * float(np.sqrt(P.Rmax2)); P.centre_ISM = (0.0, 0.0, 0.0)
    S.repartition_energie(P)
    assert np.allclose(P.frac_E_disk[7] - 0.0, P.frac_E_stars[7]) and P.frac_E_stars[7] == pytest.approx(2.0 / 3.0)
    O = Oracle(P)
    t = O.run(n_threads=1, n_photons2=200)
    n = t.stats[0]
    n_ism = n - t.stats[5] - t.stats[6]            # alive at exit but not detected
    assert 0.15 * n < n_ism < 0.34 * n             # 1/3 of the packets start as ISM; those absorbed on the way are re-emitted as disk packets
    assert t.sed.sum() == pytest.approx(t.stats[6])
    base = small_problems()["cyl2D"]()
    t0 = Oracle(base).run(n_threads=1, n_photons2=200)
    # the outer disk is heated from outside: more absorbed energy per stellar packet than without the ISM field
    outer = base.r_grid > 0.5 * base.r_grid.max()
    assert t.xKJ_abs[outer].sum() / t.stats[6] > 1.2 * t0.xKJ_abs[outer].sum() / t0.stats[6]


def test_compute_column_properties():
    """compute_column (optical_depth.f90:328-415) of the oracle: the two vertical directions add up to the same column for
    every cell of a radial ring, direction 1 equals optical_length_tot towards the origin, unit weights give path lengths."""
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    cx, cy, cz = S.cell_centres(P)
    lam = P.lambda_seuil
    col = O.compute_column(lam, cx, cy, cz).astype(np.float64)
    assert col.shape == (P.n_cells, 4)
    vert = (col[:, 1] + col[:, 2]).reshape(P.nz, P.n_rad)              # (j, i)
    assert np.allclose(vert, vert[0][None, :], rtol=2e-6)
    n = np.sqrt(cx ** 2 + cy ** 2 + cz ** 2)
    t = O.optical_length_tot(lam, cx, cy, cz, -cx / n, -cy / n, -cz / n, np.arange(1, P.n_cells + 1))
    assert np.allclose(col[:, 0], t["tau_tot"], rtol=1e-6, atol=1e-30)
    ones = O.compute_column(1, cx, cy, cz, np.ones(P.n_cells)).astype(np.float64)
    assert (ones[:, 3] > 0).all() and np.allclose(ones[:, 3] + np.hypot(cx, cy), P.r_lim[-1], rtol=1e-6)      # radial path to the outer edge
    assert np.allclose((ones[:, 1] + ones[:, 2]).reshape(P.nz, P.n_rad), 2.0 * P.zmax[None, :], rtol=1e-6)           # full height of the column


def test_define_dark_zone_oracle_matches_the_numpy_restatement():
    """The oracle's define_dark_zone (optical_depth.f90:1425-1651, nested loops as in the Fortran) against the generator's
    numpy version driven by the oracle's own ray walker (2D cylindrical): same dark cells."""
    P = S.ref41_like(n_photons_eq_th=10, dark_zone=False)
    O = Oracle(P)
    d = O.define_dark_zone(P.lambda_seuil, 1500.0, P.r_grid, P.z_grid, [(1, P.n_rad)])
    ref = S.define_dark_zone(P, P.lambda_seuil, 1500.0, Oracle(P).dark_zone_walker())
    assert d["l_dark_zone"].sum() > 100 and np.array_equal(d["l_dark_zone"], ref)
    assert d["l_is_dark_zone"] == 1 and 1 < d["ri_in"][0] < d["ri_out"][0] < P.n_rad
    dark2d = d["l_dark_zone"].reshape(P.nz, P.n_rad)
    assert (np.diff(dark2d, axis=0) <= 0).all()          # a column is dark from the midplane up to one row
    assert dark2d[:, 0].sum() == 0 and dark2d[:, -1].sum() == 0      # region edges


def test_init_reemission_oracle_matches_the_generators_tables():
    """The oracle's init_reemission (thermal_emission.f90:404-618, with the reference's `real` constants) against the tables
    the generators build in numpy with double-precision constants: equal to the 1e-8 the `1.e-6` literal is worth."""
    P = S.ref41_multi_like(n_photons_eq_th=10)
    O = Oracle(P)
    logQ, cdf = O.init_reemission(P.tab_lambda, P.tab_delta_lambda)
    ref = P.log_Qcool_minus_extra_heating
    assert np.array_equal(logQ == -1000.0, ref == -1000.0)
    ok = ref > -999.0
    assert np.abs(logQ[ok] - ref[ok]).max() < 1e-7 and np.abs(cdf - P.kdB_dT_CDF).max() < 1e-7
    assert (np.diff(cdf, axis=0) >= 0).all() and np.allclose(cdf[-1][cdf[-1] > 0], 1.0)
    Pg = S.multi_grain_like(n_photons_eq_th=10)
    Og = Oracle(Pg)
    for (k0, k1, lE, cd) in ((Pg.grain_RE_nLTE_start, Pg.grain_RE_nLTE_end, Pg.log_E_em_1grain, Pg.kdB_dT_1grain_nLTE_CDF),
                             (Pg.grain_nRE_start, Pg.grain_nRE_end, Pg.log_E_em_1grain_nRE, Pg.kdB_dT_1grain_nRE_CDF)):
        logE, Eem, c = Og.init_reemission_grains(Pg.tab_lambda, Pg.tab_delta_lambda, Pg.C_abs_norm, k0, k1)
        assert logE.shape == np.asarray(lE).shape and c.shape == np.asarray(cd).shape
        assert np.abs(logE - lE).max() < 1e-6 and np.abs(c - cd).max() < 1e-6
        assert np.allclose(np.log(Eem), logE, rtol=0, atol=1e-12) and (c[0] == 0).all()


def test_integ_ray_dust_telescopes_for_a_uniform_source_function():
    """integ_ray_dust (optical_depth.f90:1327-1421) of the oracle: with eps_dust1 = 1 everywhere the sum of
    exp(-tau)(1 - exp(-dtau)) telescopes to 1 - exp(-tau_tot), tau_tot from optical_length_tot; with tau_dark_zone_obs
    small the integration stops early; init_dust_source_fct1 without scattered light gives J_th / kappa."""
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 2000, seed=5)
    lam = 20
    eps = np.ones((45, 2, 1, P.n_cells), order="F")
    I = O.integ_ray_dust(lam, x, y, z, u, v, w, ic, 1.0e30, eps)
    tau = O.optical_length_tot(lam, x, y, z, u, v, w, ic)["tau_tot"]
    assert I.shape == (1, 2000) and np.allclose(I[0], 1.0 - np.exp(-tau), rtol=1e-6, atol=1e-7)
    I2 = O.integ_ray_dust(lam, x, y, z, u, v, w, ic, 0.5, eps)
    deep = tau > 5.0
    assert deep.sum() > 10 and (I2[0][deep] < I[0][deep]).all() and (I2[0][deep] > 1.0 - np.exp(-0.5) - 1e-12).all()
    J = np.random.default_rng(1).uniform(1.0, 2.0, P.n_cells)
    e = O.init_dust_source_fct1(lam, 1, 1, 1.0, J, np.zeros((45, 2, 1, 1, P.n_cells), np.float32), 1, False, False)
    kap = P.kappa.reshape(P.p_n_cells, P.n_lambda)[0, lam - 1] * P.kappa_factor
    ok = kap > 0
    assert np.allclose(e[3, 1, 0][ok], (J / np.where(ok, kap, 1.0))[ok], rtol=1e-14) and (e[:, :, 0][:, :, ~ok] == 0).all()


def test_repartition_energie_oracle_matches_the_generator():
    """The oracle's repartition_energie (thermal_emission.f90:1771-1949, LTE case) against the generator's numpy version
    (double-precision constants: the 1e-8 the `1.e-6` literal is worth on the wavelength, up to 1e-6 on the Wien side where
    hc / (k T lambda) ~ 80), with dark cells and cold cells."""
    P = S.ref41_like(n_photons_eq_th=10, dark_zone=False, n_rad=30, nz=16, n_rad_in=5)
    T = np.random.default_rng(1).uniform(20.0, 800.0, P.n_cells).astype(np.float32)
    T[::13] = 0.0
    dark = np.zeros(P.n_cells, np.int32); dark[3::7] = 1
    S.repartition_energie(P, Tdust=T)                      # (no dark zone)
    O = Oracle(P)
    r0 = O.repartition_energie(T, P.tab_lambda, P.E_stars)
    ok = P.E_disk > 0
    assert ok.sum() > 20 and np.allclose(r0["E_disk"][ok], P.E_disk[ok], rtol=1e-5)
    assert np.abs(r0["prob_E_cell"] - P.prob_E_cell).max() < 1e-6 and np.allclose(r0["frac_E_stars"], P.frac_E_stars, rtol=1e-5)
    O.set_dark_zone(dark)
    r = O.repartition_energie(T, P.tab_lambda, P.E_stars)
    inc = np.diff(r["prob_E_cell"], axis=0)
    assert (inc[dark == 1] == 0).all() and (inc[T == 0] == 0).all() and (r["E_disk"][ok] < r0["E_disk"][ok]).all()
    w = np.linspace(0.5, 1.5, P.n_cells)
    rw = O.repartition_energie(T, P.tab_lambda, P.E_stars, weight=w, lambda_first=10, lambda_last=12)
    assert np.allclose(rw["E_disk"][9:12], r["E_disk"][9:12]) and (rw["E_disk"][:9] == 0).all() and np.all(rw["weight_norm"][9:12] > 0)
