"""CPU tests of the modified-random-walk pieces of the oracle (MRW.f90, distance_to_closest_wall_*, mean opacities):
the reference holds nothing to compare them with (the call site is commented out, dust_transfer.f90:1222-1239), so they
are pinned by their defining properties."""
import numpy as np
import pytest

from mcfost_b200 import synthetic as S
from oracle.binding import Oracle

from helpers import rays_in_cells, small_problems


@pytest.fixture(scope="module")
def oracles():
    return {name: (P, Oracle(P)) for name, P in ((n, f()) for n, f in small_problems().items())}


def test_zeta_table_is_the_series_of_min_et_al_eq7(oracles):
    O = oracles["cyl2D"][1]
    z = O.zeta_table()
    y = np.arange(10000) / 9999.0
    assert z[0] == 0.0 and abs(z[-1] - 1.0) < 1e-14 and (np.diff(z) >= 0).all()
    for i in (1, 10, 500, 3000, 6000, 8000):
        n = np.arange(1, 400)
        assert z[i] == pytest.approx(2.0 * np.sum((-1.0) ** (n + 1) * y[i] ** (n.astype(float) ** 2)), rel=1e-13, abs=1e-300)
    # small y: zeta = 2 y; close to 1 the series saturates (1 - zeta = 2 sqrt(pi/eps) exp(-pi^2 / (4 eps)), eps = -ln y)
    assert z[3] == pytest.approx(2 * y[3], rel=1e-10)
    eps = -np.log(y[7000])
    assert 1.0 - z[7000] == pytest.approx(2 * np.sqrt(np.pi / eps) * np.exp(-np.pi ** 2 / (4 * eps)), rel=1e-3)
    # sample_zeta inverts the table (interp of utils.f90:190-247)
    for zr in (1e-7, 1e-3, 0.2, 0.5, 0.9, 0.999, 1 - 2.0 ** -24):
        ym = O.sample_zeta(zr)
        assert 0.0 <= ym < 1.0
        assert np.interp(ym, y, z) == pytest.approx(zr, rel=1e-9)


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D", "sph3D"])
def test_sphere_of_radius_closest_wall_stays_in_the_cell(oracles, name):
    P, O = oracles[name]
    ic, x, y, z, *_ = rays_in_cells(P, 4000, seed=3)
    d = O.distance_to_closest_wall(ic, x, y, z)
    assert (d > 0).all()
    rng = np.random.default_rng(5)
    for _ in range(6):
        w = rng.uniform(-1, 1, len(x)); ph = rng.uniform(0, 2 * np.pi, len(x))
        u, v = np.sqrt(1 - w * w) * np.cos(ph), np.sqrt(1 - w * w) * np.sin(ph)
        f = d * (1 - 1e-9)
        inside = O.index_cell(x + f * u, y + f * v, z + f * w)
        if not P.l3D:      # 2D grids number a cell and its mirror image below the midplane identically
            assert np.array_equal(inside, ic)
        else:
            assert np.array_equal(inside, ic)
    # it is the distance to the CLOSEST wall: some direction leaves the cell just beyond it (cylindrical: exact walls)
    if name == "cyl2D":
        ci, cj = P.cell_map_i[ic - 1], P.cell_map_j[ic - 1]
        r = np.sqrt(x * x + y * y)
        ref = np.minimum.reduce([P.r_lim[ci] - r, r - P.r_lim[ci - 1], P.z_lim[ci - 1, cj] - np.abs(z), np.abs(z) - P.z_lim[ci - 1, cj - 1]])
        assert np.allclose(d, ref, rtol=1e-12, atol=0)


def test_voronoi_closest_wall_is_a_length(oracles):
    P = S.voronoi_disk(n_points=600, n_photons_eq_th=10)
    O = Oracle(P)
    rng = np.random.default_rng(2)
    ic = rng.integers(1, P.n_cells + 1, 500).astype(np.int32)
    x, y, z = (P.vor_xyz[a, ic - 1].copy() for a in range(3))
    d = O.distance_to_closest_wall(ic, x, y, z)
    # from the seed itself the closest face is half the distance to the nearest neighbour (fp32 geometry)
    cut = np.asarray(P.vor_was_cut)[ic - 1] != 0
    for q in np.flatnonzero(~cut)[:100]:
        i = ic[q]
        nb = P.neighbours_list[P.vor_first[i - 1] - 1:P.vor_last[i - 1]]
        if (nb < 0).any():
            assert d[q] == 0.0
            continue
        dist = np.sqrt(((P.vor_xyz[:, nb - 1] - P.vor_xyz[:, [i - 1]]) ** 2).sum(axis=0))
        assert d[q] == pytest.approx(0.5 * dist.min(), rel=1e-4)
    assert (d[cut] == 0).all()


def test_mean_opacities_are_the_moments_of_the_reemission_spectrum(oracles):
    P, O = oracles["cyl2D"]
    A, B, Cc = O.mrw_tables()
    cdf = np.asarray(P.kdB_dT_CDF).reshape(P.n_lambda, P.n_T, P.p_n_cells, order="F")
    p = np.diff(np.concatenate([np.zeros((1, P.n_T, P.p_n_cells)), cdf]), axis=0)
    kap = np.asarray(P.kappa).reshape(P.p_n_cells, P.n_lambda, order="F").T[:, None, :]
    alb = np.asarray(P.tab_albedo_pos, np.float64).reshape(P.p_n_cells, P.n_lambda, order="F").T[:, None, :]
    g = np.asarray(P.tab_g_pos, np.float64).reshape(P.p_n_cells, P.n_lambda, order="F").T[:, None, :]
    kabs = np.asarray(P.kappa_abs_LTE).reshape(P.p_n_cells, P.n_lambda, order="F").T[:, None, :]
    k_a, k_t = kap * (1 - alb), kap * (1 - alb * g)
    assert np.allclose(A, (p / k_a).sum(axis=0), rtol=1e-12)
    assert np.allclose(B, (p / (k_a * k_t)).sum(axis=0), rtol=1e-12)
    assert np.allclose(Cc, (p * kabs / k_a).sum(axis=0), rtol=1e-12)
    # kappa_abs_LTE = kappa (1 - albedo) in an LTE-only model: one absorption per cycle
    assert np.allclose(Cc[1:], 1.0, rtol=1e-5)
    # grey limit: the Rosseland-type mean free path is 1 / kappa_transport
    ratio = B / A
    assert (ratio > (1 / k_t).min() * 0.999).all() and (ratio < (1 / k_t).max() * 1.001).all()


def test_mrw_leaves_the_temperature_unchanged():
    """MRW on and off on an optically thick disk, same packets: the temperatures agree well inside the Monte Carlo noise
    of two independent runs, the emergent spectrum too, and the walk replaces a large share of the interactions."""
    P = S.ref41_like(n_photons_eq_th=1500, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e4)
    O = Oracle(P, fast=True)
    P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, O.dark_zone_walker())
    S.repartition_energie(P)
    O.set_dark_zone(P.l_dark_zone); O.set_emission(P)
    t0 = O.run(n_threads=0, n_photons2=1500, lMRW=0); T0 = O.temp_finale()
    t1 = O.run(n_threads=0, n_photons2=1500, lMRW=1); T1 = O.temp_finale()
    t2 = O.run(n_threads=0, n_photons2=1500, lMRW=0, seed=77); T2 = O.temp_finale()
    assert t1.stats[8] > 0 and t1.stats[9] >= t1.stats[8]
    assert t1.stats[2] < 0.97 * t0.stats[2]                      # fewer interactions
    assert t0.stats[8] == 0 and t0.stats[9] == 0
    for t in (t0, t1):
        assert t.stats[5] + t.stats[6] == t.stats[0] == 128 * 1500
    lit = (T0 > 1.5) & (np.asarray(P.l_dark_zone) == 0) & (t0.xKJ_abs > 0) & (t2.xKJ_abs > 0)
    rel_mrw = np.abs(T1[lit] - T0[lit]) / T0[lit]
    rel_seed = np.abs(T2[lit] - T0[lit]) / T0[lit]
    assert np.median(rel_mrw) < 0.01 and np.percentile(rel_mrw, 75) < 0.05          # the reference's MC_similar bars
    assert np.median(rel_mrw) < 1.5 * np.median(rel_seed)
    assert abs(np.mean((T1[lit] - T0[lit]) / T0[lit])) < 0.01                        # no bias
    # emergent spectrum (packets per wavelength, all inclinations)
    n0, n1 = t0.n_phot_sed.sum(axis=(1, 2)), t1.n_phot_sed.sum(axis=(1, 2))
    m = (n0 + n1) > 100
    z = (n1[m] - n0[m]) / np.sqrt(n0[m] + n1[m])
    assert np.mean(np.abs(z) < 3) > 0.95 and abs(z.mean()) < 0.5
    # lMRW is a thermal-step, LTE-only option
    with pytest.raises(RuntimeError):
        O.run(n_threads=1, n_photons2=1, lMRW=1, letape_th=0, lmono=1)
