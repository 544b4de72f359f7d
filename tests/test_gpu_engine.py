"""GPU tests (-m gpu) of the round-2 pieces of the thermal step, all through the C ABI, at the PRODUCTION launch
configuration (no tuning knobs): the packet-per-warp kernel that runs small budgets and the last packets of large
ones, the three-launch hand-over, distance_to_closest_wall, the modified random walk and its tables."""
import numpy as np
import pytest

from mcfost_b200 import api, synthetic as S
from oracle.binding import Oracle

from helpers import rays_in_cells, small_problems

pytestmark = pytest.mark.gpu

GRIDS = ["cyl2D", "cyl3D", "sph2D", "sph3D"]


def _spectrum_z(ta, tb, min_count=100):
    """z-scores of the emergent packet counts per wavelength (all inclinations) of two independent runs."""
    na, nb = ta.n_phot_sed.sum(axis=(1, 2)), tb.n_phot_sed.sum(axis=(1, 2))
    m = (na + nb) > min_count
    return (na[m] - nb[m]) / np.sqrt(na[m] + nb[m])


def _energy_spectrum_close(ta, tb):
    """energy-weighted emergent spectrum: sed per wavelength within 3 sigma of its own packet statistics"""
    ea, eb = ta.sed.sum(axis=(1, 2)), tb.sed.sum(axis=(1, 2))
    na, nb = ta.n_phot_sed.sum(axis=(1, 2)), tb.n_phot_sed.sum(axis=(1, 2))
    m = (na + nb) > 100
    sig = np.sqrt(ea[m] ** 2 / np.maximum(na[m], 1) + eb[m] ** 2 / np.maximum(nb[m], 1))
    return np.abs(ea[m] - eb[m]) / sig


@pytest.mark.parametrize("name", GRIDS)
def test_distance_to_closest_wall_bit_exact(name):
    P = small_problems()[name]()
    O, G = Oracle(P), api.PhotonLoop(P)
    ic, x, y, z, *_ = rays_in_cells(P, 100000, seed=31)
    assert np.array_equal(G.distance_to_closest_wall(ic, x, y, z), O.distance_to_closest_wall(ic, x, y, z))
    with pytest.raises(api.McfostB200Error):
        G.distance_to_closest_wall(np.array([P.n_cells + 1], np.int32), x[:1], y[:1], z[:1])
    G.close()


def test_distance_to_closest_wall_voronoi_bit_exact():
    P = S.voronoi_disk(n_points=1500, n_photons_eq_th=300)
    O, G = Oracle(P), api.PhotonLoop(P)
    rng = np.random.default_rng(4)
    ic = rng.integers(1, P.n_cells + 1, 20000).astype(np.int32)
    jit = 1e-3 * rng.normal(size=(3, len(ic))) * np.asarray(P.vor_h)[ic - 1]
    x, y, z = (P.vor_xyz[a, ic - 1] + jit[a] for a in range(3))
    assert np.array_equal(G.distance_to_closest_wall(ic, x, y, z), O.distance_to_closest_wall(ic, x, y, z))
    G.close()


def test_mrw_mean_opacity_tables_match_oracle():
    for P in (small_problems()["cyl2D"](), S.ref41_multi_like(n_photons_eq_th=10)):
        O, G = Oracle(P), api.PhotonLoop(P)
        for a, b in zip(G.mrw_tables(), O.mrw_tables()):
            assert np.array_equal(a, b)
        G.close()


@pytest.mark.parametrize("name", GRIDS)
def test_para_budget_spectrum_every_grid(name):
    """128 x 1000 packets (the budget of the reference's own .para files): the packet-per-warp kernel runs the whole
    call.  Emergent spectrum per wavelength (counts AND energy) against the oracle within 3 sigma, temperatures within the
    north_star bars."""
    P = small_problems()[name]()
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 1000, 1.0e30, 1, False)
    Tg = G.temp_finale()
    assert G.debug_counters()["parked"] == 0
    G.close()
    O = Oracle(P, fast=True)
    to = O.run(n_threads=0, n_photons2=1000)
    To = O.temp_finale()
    assert tg.stats[0] == to.stats[0] == 128000 == tg.n_phot_envoyes.sum()
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]
    assert tg.sed.sum() == pytest.approx(tg.stats[6] * P.E_paquet)
    z = _spectrum_z(tg, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5, z
    assert np.mean(_energy_spectrum_close(tg, to) < 3) >= 0.95
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0) & (To > 1.5)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert np.median(rel) < 0.02 and np.percentile(rel, 75) < 0.05, (np.median(rel), np.percentile(rel, 75))
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.03
    assert abs(tg.stats[1] / to.stats[1] - 1) < 0.02 and abs(tg.stats[2] / to.stats[2] - 1) < 0.03


def test_para_budget_with_stokes_and_variable_dust():
    """the same for the polarised call of ref4.1 (lsepar_pola) and for cell-dependent tables (global-memory kernels)"""
    P = small_problems()["cyl2D"]()
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 1000, 1.0e30, 1, False, lsepar_pola=1, lsepar_contrib=1)
    G.close()
    to = Oracle(P, fast=True).run(n_threads=0, n_photons2=1000, lsepar_pola=1, lsepar_contrib=1)
    z = _spectrum_z(tg, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5
    qg, qo = np.abs(tg.sed_q).sum(), np.abs(to.sed_q).sum()
    assert qg > 0 and abs(qg / qo - 1) < 0.1
    assert abs(tg.sed_star_scat.sum() / to.sed_star_scat.sum() - 1) < 0.05
    P = S.ref41_multi_like(n_photons_eq_th=1000)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 1000, 1.0e30, 1, False)
    G.close()
    to = Oracle(P, fast=True).run(n_threads=0, n_photons2=1000)
    z = _spectrum_z(tg, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.03


def test_three_launch_call_matches_oracle():
    """2.56e6 packets on the G1 geometry at test size: packet-per-warp kernel first, packet-per-lane kernel for the bulk
    with the packets in flight capped at a fraction of those sent, stragglers back on the packet-per-warp kernel."""
    P = S.ref41_like(n_photons_eq_th=20000, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 20000, 1.0e30, 1, False, lsepar_pola=1)
    Tg = G.temp_finale()
    d = G.debug_counters()
    G.close()
    O = Oracle(P, fast=True)
    to = O.run(n_threads=0, n_photons2=20000, lsepar_pola=1)
    To = O.temp_finale()
    assert d["parked"] > 0
    assert tg.stats[0] == to.stats[0] == 128 * 20000 == tg.n_phot_envoyes.sum()
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]
    assert tg.sed.sum() == pytest.approx(tg.stats[6])
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0) & (To > 1.5)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05, (np.median(rel), np.percentile(rel, 75))
    z = _spectrum_z(tg, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5, z
    assert abs(tg.stats[1] / to.stats[1] - 1) < 0.01 and abs(tg.stats[2] / to.stats[2] - 1) < 0.01
    assert abs(np.abs(tg.sed_q).sum() / np.abs(to.sed_q).sum() - 1) < 0.05


@pytest.mark.parametrize("n2", [1500, 12000])
def test_modified_random_walk_matches_oracle_and_plain_run(n2):
    """lMRW on an optically thick disk, small budget (packet-per-warp kernel: every eligible packet walks) and three-launch
    budget (the packet-per-lane kernel skips the walk, which is always allowed): same physics as the oracle's MRW and as
    the run without it."""
    P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e4)
    O = Oracle(P, fast=True)
    P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, O.dark_zone_walker())
    S.repartition_energie(P)
    O.set_dark_zone(P.l_dark_zone); O.set_emission(P)
    G = api.PhotonLoop(P)
    t0 = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False); T0 = G.temp_finale()
    t1 = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False, lMRW=1); T1 = G.temp_finale()
    with pytest.raises(api.McfostB200Error):
        G.mc_photon_loop(8, 8, 10, 64.0, 1, False, letape_th=0, lmono=1, lMRW=1)
    G.close()
    to = O.run(n_threads=0, n_photons2=n2, lMRW=1); To = O.temp_finale()
    assert t0.stats[8] == 0 and t1.stats[8] > 0 and t1.stats[9] >= t1.stats[8]
    for t in (t0, t1):
        assert t.stats[0] == 128 * n2 and t.stats[5] + t.stats[6] == t.stats[0]
    if 128 * n2 <= 1000000:      # the packet-per-warp kernel runs the whole call: walks and steps per packet like the oracle's
        # (seed-to-seed spread of the oracle at this budget: interactions 2.5 %, walks 5.4 %, walk steps 6.3 % -- a handful of
        # trapped packets make most of them; the ratio of interactions with / without the walk is 0.87 +- 0.02)
        assert t1.stats[2] < 0.97 * t0.stats[2]
        assert abs(t1.stats[8] / to.stats[8] - 1) < 0.2 and abs(t1.stats[9] / to.stats[9] - 1) < 0.25
        assert abs(t1.stats[2] / to.stats[2] - 1) < 0.1
    else:                        # three launches: only the packets the packet-per-warp kernel finishes walk (MRW)
        assert t1.stats[8] < to.stats[8]
    lit = (np.asarray(P.l_dark_zone) == 0) & (t0.xKJ_abs > 0) & (t1.xKJ_abs > 0) & (to.xKJ_abs > 0) & (To > 1.5)
    for Ta, Tb in ((T1, To), (T1, T0)):
        rel = np.abs(Ta[lit] - Tb[lit]) / Tb[lit]
        assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05, (np.median(rel), np.percentile(rel, 75))
        assert abs(np.mean((Ta[lit] - Tb[lit]) / Tb[lit])) < 0.01
    z = _spectrum_z(t1, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5
    z = _spectrum_z(t1, t0)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5


def test_modified_random_walk_other_grids():
    for name in ("cyl3D", "sph2D"):
        P = small_problems()[name]()
        G = api.PhotonLoop(P)
        t0 = G.mc_photon_loop(1, 1, 1000, 1.0e30, 1, False); T0 = G.temp_finale()
        t1 = G.mc_photon_loop(1, 1, 1000, 1.0e30, 1, False, lMRW=1); T1 = G.temp_finale()
        G.close()
        assert t1.stats[5] + t1.stats[6] == t1.stats[0] == 128000
        lit = (t0.xKJ_abs > 0) & (t1.xKJ_abs > 0) & (T0 > 1.5)
        rel = np.abs(T1[lit] - T0[lit]) / T0[lit]
        assert np.median(rel) < 0.02 and np.percentile(rel, 75) < 0.05


@pytest.mark.parametrize("tau_V", [0.1, 1.0, 10.0, 100.0])
def test_pascucci_benchmark_matches_oracle(tau_V):
    """G2: Pascucci_3.0.para (2D disk benchmark, single 0.12 um grain, isotropic scattering, 61 wavelengths) at the
    file's own budget of 1.28e6 thermal packets, for the four optical depths of the benchmark."""
    n2 = 10000
    P = S.pascucci_like(tau_V=tau_V, n_photons_eq_th=n2)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False, lisotropic=1)
    Tg = G.temp_finale()
    G.close()
    O = Oracle(P, fast=True)
    to = O.run(n_threads=0, n_photons2=n2, lisotropic=1)
    To = O.temp_finale()
    assert tg.stats[0] == to.stats[0] == 128 * n2 == tg.n_phot_envoyes.sum()
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0) & (To > 1.5)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert lit.sum() > 0.8 * P.n_cells
    assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05, (np.median(rel), np.percentile(rel, 75))
    z = _spectrum_z(tg, to)
    assert np.mean(np.abs(z) < 3) >= 0.95 and abs(z.mean()) < 0.5, z
    assert np.mean(_energy_spectrum_close(tg, to) < 3) >= 0.95
    assert abs(tg.stats[1] / to.stats[1] - 1) < 0.01
    assert abs(tg.stats[2] - to.stats[2]) < 5 * np.sqrt(to.stats[2]) + 0.01 * to.stats[2]


def test_voronoi_point_location_grid_equals_brute_force():
    """GeomVor::index walks a uniform grid of seeds instead of the reference's O(n_cells) scan (index_cell_voronoi,
    Voronoi.f90:1548-1572): same nearest seed (fp32 distances, ties to the lowest id) for 1e5 points, including points
    on seeds, midway between seeds and outside the seeds' bounding box."""
    P = S.voronoi_disk(n_points=1500, n_photons_eq_th=10)
    O, G = Oracle(P), api.PhotonLoop(P)
    rng = np.random.default_rng(11)
    n = 100000
    L = 100.0
    x, y, z = (rng.uniform(-L, L, n) for _ in range(3))
    # on seeds, midpoints of random seed pairs, far outside
    s = rng.integers(0, P.n_cells, 3000)
    x[:1000], y[:1000], z[:1000] = P.vor_xyz[0, s[:1000]], P.vor_xyz[1, s[:1000]], P.vor_xyz[2, s[:1000]]
    a, b = s[1000:2000], s[2000:3000]
    x[1000:2000], y[1000:2000], z[1000:2000] = (0.5 * (P.vor_xyz[k, a] + P.vor_xyz[k, b]) for k in range(3))
    x[2000:2500] *= 5.0; y[2000:2500] *= 5.0; z[2500:3000] *= 7.0
    assert np.array_equal(G.index_cell(x, y, z), O.index_cell(x, y, z))
    G.close()


def _n_cuda_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("n_gpus", [1, 2])
def test_multi_gpu_behind_the_abi_equals_one_gpu(n_gpus):
    """mcfost_b200_multi_*: n GPUs, one call, every tally merged inside the library.  The SED step has no feedback, so
    the union of the ranks' packets is exactly the single-GPU set: counts identical, sums to rounding; xI_scatt too.
    The thermal step (feedback: local tally x n) agrees statistically, and Temp_finale runs on the merged tallies."""
    if _n_cuda_devices() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, lscatt_ray_tracing1=1, lsepar_pola=1, lsepar_contrib=1, RT_n_incl=3, RT_n_az=1,
              tab_u_rt=np.array([[0.0], [0.5], [0.9]]), tab_v_rt=np.zeros((3, 1)), tab_w_rt=np.array([1.0, np.sqrt(0.75), np.sqrt(0.19)]))
    G1 = api.PhotonLoop(P)
    a = G1.mc_photon_loop(8, 8, 10 ** 9, 200.0, 1, False, **kw)
    th1 = G1.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False); T1 = G1.temp_finale()
    G1.close()
    M = api.MultiPhotonLoop(P, n_gpus)
    b = M.mc_photon_loop(8, 8, 10 ** 9, 200.0, 1, False, **kw)
    thn = M.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False); Tn = M.temp_finale()
    M.close()
    assert b.stats[0] == a.stats[0] == 128 * 200
    assert np.array_equal(b.n_phot_sed, a.n_phot_sed) and np.array_equal(b.n_phot_envoyes, a.n_phot_envoyes)
    assert np.array_equal(b.stats[:8], a.stats[:8])
    for k in ("sed", "sed_q", "sed_u", "sed_star_scat", "sed_disk_scat"):
        assert np.allclose(getattr(b, k), getattr(a, k), rtol=1e-9, atol=1e-12), k
    assert np.allclose(b.xI_scatt, a.xI_scatt, rtol=2e-4, atol=1e-6 * np.abs(a.xI_scatt).max())      # fp32 sums, different order
    assert thn.stats[0] == th1.stats[0] == 128 * 1500 and thn.stats[5] + thn.stats[6] == thn.stats[0]
    assert abs(thn.xKJ_abs.sum() / th1.xKJ_abs.sum() - 1) < 0.02
    lit = (th1.xKJ_abs > 0) & (thn.xKJ_abs > 0) & (T1 > 1.5)
    assert np.median(np.abs(Tn[lit] - T1[lit]) / T1[lit]) < 0.02
    assert (thn.xT_ech >= 2).all()


@pytest.mark.parametrize("name", ["cyl2D", "sph3D"])
def test_device_temp_finale_equals_the_oracles_on_the_same_tallies(name):
    """mcfost_b200_temp_finale against the ORACLE's Temp_finale (thermal_emission.f90:870-906, oracle.cpp) fed with the very
    tallies the device holds: same table walk, `real` output equal to libm rounding."""
    P = small_problems()[name]()
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 500, 1.0e30, 1, False)
    Tg = G.temp_finale()
    G.close()
    To = Oracle(P).temp_finale_of(tg.xKJ_abs, tg.xT_ech)
    assert np.allclose(Tg, To, rtol=5e-6, atol=0)
    assert (Tg > P.T_min).sum() > 0.5 * P.n_cells


def test_interstellar_side_loop_matches_oracle():
    """lISM_loop: the ISM side loop of run_sed_mc (dust_transfer.f90:941-985, lProDiMo / lML): every packet from
    emit_packet_ISM, chunks end when n_photons2 packets ENTERED the model, n_phot_envoyes counts all of them, xJ_abs is the
    tally of interest.  The device lets the packets in flight complete when a chunk is full (a few hundred more packets per
    chunk than the oracle, whatever n_photons2), so the comparison is per packet sent."""
    P = S.ref41_like(n_photons_eq_th=100, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3)
    P.R_ISM = 1.2 * np.sqrt(P.Rmax2 + P.zmaxmax ** 2); P.centre_ISM = (0.0, 0.0, 0.0)
    kw = dict(letape_th=0, lmono=1, lISM_loop=1, lxJ_abs=1, lambda_in=20, p_lambda_in=20, n_photons2=4000, n_phot_lim=1.0e30)
    G = api.PhotonLoop(P)
    k2 = dict(kw)
    tg = G.mc_photon_loop(k2.pop("lambda_in"), k2.pop("p_lambda_in"), k2.pop("n_photons2"), k2.pop("n_phot_lim"), 1, False, **k2)
    with pytest.raises(api.McfostB200Error):
        G.mc_photon_loop(1, 1, 10, lISM_loop=1)       # not a thermal-step mode
    G.close()
    to = Oracle(P).run(n_threads=0, xJ=True, **kw)
    lam = 20
    assert to.n_phot_envoyes[lam - 1] == to.stats[0] and tg.n_phot_envoyes[lam - 1] == tg.stats[0]
    assert tg.sed.sum() == 0 and to.sed.sum() == 0                          # ISM packets are never detected (:548)
    assert tg.stats[0] >= 128 * 4000 and to.stats[0] >= 128 * 4000           # packets that miss the model are sent again
    assert 0.99 < tg.stats[0] / to.stats[0] < 1.15
    jg, jo = tg.xJ_abs[:, lam - 1].sum() / tg.stats[0], to.xJ_abs[:, lam - 1].sum() / to.stats[0]
    assert abs(jg / jo - 1) < 0.03
    assert abs(tg.stats[1] / tg.stats[0] / (to.stats[1] / to.stats[0]) - 1) < 0.03


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D", "sph3D", "voronoi", "variable_dust"])
def test_compute_column_bit_exact(name):
    """compute_column (optical_depth.f90:328-415): optical depth and weighted column from every cell centre along the four
    directions, `real` output identical to the oracle's (same cross_cell sequence, same summation order)."""
    if name == "voronoi":
        P = S.voronoi_disk(n_points=1500, n_photons_eq_th=10)
    elif name == "variable_dust":
        P = S.ref41_multi_like(n_photons_eq_th=10)
    else:
        P = small_problems()[name]()
    cx, cy, cz = S.cell_centres(P)
    O, G = Oracle(P), api.PhotonLoop(P)
    weight = np.random.default_rng(3).uniform(0.5, 2.0, P.n_cells)
    for lam, f in ((P.lambda_seuil, None), (1, None), (1, weight)):
        o, g = O.compute_column(lam, cx, cy, cz, f), G.compute_column(lam, cx, cy, cz, f)
        assert o.shape == g.shape == (P.n_cells, 4) and np.array_equal(o, g)
        assert np.isfinite(g).all() and (g >= 0).all() and g[:, 1:3].max() > 0
    with pytest.raises(api.McfostB200Error):
        G.compute_column(P.n_lambda + 1, cx, cy, cz)
    G.close()


@pytest.mark.parametrize("name", ["cyl2D_full", "cyl2D_small", "cyl3D", "sph2D", "variable_dust"])
def test_define_dark_zone_in_the_library_equals_oracle(name):
    """mcfost_b200_define_dark_zone (optical_depth.f90:1425-1651, one block: the columns in sequence, the rays of a column in
    parallel) against the oracle's nested loops: identical l_dark_zone, ri_in / ri_out, zj_sup / zj_inf and flag."""
    if name == "cyl2D_full":
        P = S.ref41_like(n_photons_eq_th=10, dark_zone=False); tau_max = 1500.0
    elif name == "cyl2D_small":
        P = small_problems()["cyl2D"](); tau_max = 30.0
    elif name == "cyl3D":
        P = S.ref41_3d_like(n_photons_eq_th=10, n_rad=30, nz=10, n_az=12, n_rad_in=4, tau_mid=3000.0); tau_max = 100.0
    elif name == "sph2D":
        P = S.spherical_shell(n_photons_eq_th=10, tau_mid=3000.0); tau_max = 50.0
    else:
        P = S.ref41_multi_like(n_photons_eq_th=10, tau_mid=3.0e4); tau_max = 300.0
    lam = P.lambda_seuil
    regions = [(1, P.n_rad)]
    O, G = Oracle(P), api.PhotonLoop(P)
    o = O.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions)
    g = G.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions)
    for k in ("l_dark_zone", "ri_in", "ri_out", "zj_sup", "zj_inf"):
        assert np.array_equal(o[k], g[k]), k
    assert o["l_is_dark_zone"] == g["l_is_dark_zone"]
    assert g["l_dark_zone"].sum() > 0, "the model has no dark zone: the test would be empty"
    if name == "cyl2D_full":      # the path bench.py takes (numpy logic + the library's ray walker)
        assert np.array_equal(g["l_dark_zone"], S.define_dark_zone(P, lam, tau_max, G.dark_zone_walker()))
    # the result is installed: packets bounce off the dark cells (not on the shell: its dark zone encloses the star, no
    # packet could ever leave -- in the reference as well)
    if name != "sph2D":
        t = G.mc_photon_loop(1, 1, 50, 1.0e30, 1, False)
        assert t.stats[0] == 128 * 50 and t.stats[5] + t.stats[6] == t.stats[0]
        if name == "cyl2D_full":
            assert t.stats[7] > 0          # bounces
    # dust-free cells (n_zones > 1, :1634-1638)
    ds = np.ones(P.n_cells); ds[np.flatnonzero(g["l_dark_zone"])[:3]] = 0.0
    o2 = O.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions, dust_sum=ds)
    g2 = G.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions, dust_sum=ds)
    assert np.array_equal(o2["l_dark_zone"], g2["l_dark_zone"]) and g2["l_dark_zone"].sum() == g["l_dark_zone"].sum() - 3
    # stale in-out arrays (the module arrays keep the values of the previous call where no sum reaches tau_max)
    zs = np.full_like(g["zj_sup"], 2)
    o3 = O.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions, zj_sup=zs.copy())
    g3 = G.define_dark_zone(lam, tau_max, P.r_grid, P.z_grid, regions, zj_sup=zs.copy())
    assert np.array_equal(o3["l_dark_zone"], g3["l_dark_zone"]) and np.array_equal(o3["zj_sup"], g3["zj_sup"])
    G.close()


def test_init_reemission_on_the_device_equals_oracle():
    """mcfost_b200_init_reemission / _grains (thermal_emission.f90:404-618) against the oracle: same operations in the same
    order, exp / log from CUDA instead of glibc (<= 1 ulp each): 1e-12.  The tables are installed: a thermal step run on them
    equals the step run on uploaded tables."""
    P = S.ref41_multi_like(n_photons_eq_th=300)
    O, G = Oracle(P), api.PhotonLoop(P)
    t_up = G.mc_photon_loop(1, 1, 300, 1.0e30, 1, False)
    lo, co = O.init_reemission(P.tab_lambda, P.tab_delta_lambda)
    lg, cg = G.init_reemission(P.tab_lambda, P.tab_delta_lambda)
    assert np.array_equal(lo == -1000.0, lg == -1000.0)
    assert np.allclose(lg, lo, rtol=1e-12, atol=1e-12) and np.allclose(cg, co, rtol=1e-12, atol=1e-15)
    t_dev = G.mc_photon_loop(1, 1, 300, 1.0e30, 1, False)
    assert t_dev.stats[0] == t_up.stats[0] and t_dev.stats[5] + t_dev.stats[6] == t_dev.stats[0]
    assert abs(t_dev.xKJ_abs.sum() / t_up.xKJ_abs.sum() - 1) < 0.02 and abs(t_dev.stats[4] / t_up.stats[4] - 1) < 0.05
    G.close()
    # a handle that never received the two tables from the host
    import copy
    P2 = copy.copy(P); P2.log_Qcool_minus_extra_heating = None; P2.kdB_dT_CDF = None
    G2 = api.PhotonLoop(P2)
    G2.init_reemission(P.tab_lambda, P.tab_delta_lambda, download=False)
    t2 = G2.mc_photon_loop(1, 1, 300, 1.0e30, 1, False)
    assert np.array_equal(t2.xKJ_abs, t_dev.xKJ_abs) or abs(t2.xKJ_abs.sum() / t_dev.xKJ_abs.sum() - 1) < 0.02
    G2.close()
    Pg = S.multi_grain_like(n_photons_eq_th=10)
    Og, Gg = Oracle(Pg), api.PhotonLoop(Pg)
    for (k0, k1) in ((Pg.grain_RE_nLTE_start, Pg.grain_RE_nLTE_end), (Pg.grain_nRE_start, Pg.grain_nRE_end), (Pg.grain_RE_LTE_start, Pg.grain_RE_LTE_end)):
        eo = Og.init_reemission_grains(Pg.tab_lambda, Pg.tab_delta_lambda, Pg.C_abs_norm, k0, k1)
        eg = Gg.init_reemission_grains(Pg.tab_lambda, Pg.tab_delta_lambda, Pg.C_abs_norm, k0, k1)
        for a, b in zip(eo, eg):
            assert a.shape == b.shape and np.allclose(b, a, rtol=1e-12, atol=1e-15)
    Gg.close()


@pytest.mark.parametrize("n_stars", [3, 12])
def test_several_stars_packet_by_packet(n_stars):
    """Several stars (stars.f90:581-605 CDF_E_star, emit_packet :select_etoile, intersect_stars :812-884): a compact multiple
    system inside the inner cavity.  Forced-scattering mode has no feedback, so with shared Philox streams the GPU follows the
    oracle's packets; then a thermal step, statistically.  12 stars: above the 8 of round 1 (15 fit the packed state)."""
    P = small_problems()["cyl2D"]()
    rng = np.random.default_rng(7)
    P.n_stars = n_stars
    xyz = rng.uniform(-0.35, 0.35, (3, n_stars)); xyz[2] *= 0.3; xyz[:, 0] = 0.0
    rad = rng.uniform(1.0, 3.0, n_stars) * S.RSUN_TO_AU
    P.star_xyzr = np.asfortranarray(np.vstack([xyz, rad[None, :]]))
    P.star_T = rng.uniform(3500.0, 9000.0, n_stars); P.star_out_model = np.zeros(n_stars, np.int32)
    S.locate_stars(P, Oracle(P).index_cell)
    S.star_energy(P); S.repartition_energie(P)
    assert P.CDF_E_star.shape == (P.n_lambda, n_stars + 1)
    from test_gpu_parity import _sed_kwargs
    kw = _sed_kwargs(False)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(kw.pop("lambda_in"), kw.pop("p_lambda_in"), kw.pop("n_photons2"), kw.pop("n_phot_lim"), 1, False, **kw)
    to = Oracle(P).run(n_threads=0, n_xI=45 * 2 * 5 * 3 * P.n_cells, **_sed_kwargs(False))
    assert tg.stats[0] == to.stats[0]
    assert abs(tg.stats[1] - to.stats[1]) <= 1e-4 * to.stats[1] and abs(tg.stats[2] - to.stats[2]) <= 1e-4 * to.stats[2]
    assert abs(tg.stats[5] - to.stats[5]) <= 3                            # packets that end on a star (or fade out)
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3) and np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=2e-3)
    assert np.allclose(tg.sed_star.sum(axis=0), to.sed_star.sum(axis=0), rtol=2e-3)
    # thermal step
    th_g = G.mc_photon_loop(1, 1, 800, 1.0e30, 1, False)
    G.close()
    th_o = Oracle(P, fast=True).run(n_threads=0, n_photons2=800)
    assert th_g.stats[0] == th_o.stats[0] == 128 * 800
    assert abs(th_g.xKJ_abs.sum() / th_o.xKJ_abs.sum() - 1) < 0.02 and abs(th_g.stats[5] / max(th_o.stats[5], 1) - 1) < 0.3
    To, Tg = S.temp_finale(P, th_o.xKJ_abs), S.temp_finale(P, th_g.xKJ_abs)
    lit = (th_o.xKJ_abs > 0) & (th_g.xKJ_abs > 0)
    assert np.median(np.abs(Tg[lit] - To[lit]) / To[lit]) < 0.02
    Pbad = small_problems()["cyl2D"]()
    Pbad.n_stars = 16
    Pbad.star_xyzr = np.asfortranarray(np.tile(Pbad.star_xyzr, (1, 16))); Pbad.star_T = np.full(16, 5000.0)
    Pbad.star_out_model = np.zeros(16, np.int32); Pbad.star_icell = np.full(16, Pbad.star_icell[0], np.int32)
    S.star_energy(Pbad)
    with pytest.raises(api.McfostB200Error):
        api.PhotonLoop(Pbad)


@pytest.mark.parametrize("pola", [False, True])
def test_rt1_source_function_and_formal_solution_match_oracle(pola):
    """Ray-tracing method 1 after an SED step: init_dust_source_fct1 (dust_ray_tracing.f90:636-708) on the xI_scatt tally that
    the step left on the device, then integ_ray_dust (optical_depth.f90:1327-1421) along random rays, against the oracle fed
    with the same tally: eps_dust1 is pure arithmetic (identical), the formal solution has exp / atan2 (1e-11)."""
    from test_gpu_parity import _sed_kwargs
    P = small_problems()["cyl2D"]()
    kw = _sed_kwargs(pola)
    lam = kw["lambda_in"]
    ntf = (4 if pola else 1) + 4
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(kw.pop("lambda_in"), kw.pop("p_lambda_in"), kw.pop("n_photons2"), kw.pop("n_phot_lim"), 1, False, **kw)
    assert tg.xI_scatt.size == 45 * 2 * ntf * 3 * P.n_cells and np.abs(tg.xI_scatt).sum() > 0
    J = np.random.default_rng(2).uniform(0.0, 1.0e-3, P.n_cells) * np.abs(tg.xI_scatt).max()
    O = Oracle(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 20000, seed=9)
    for iRT in (1, 3):
        eg = G.init_dust_source_fct1(lam, iRT, 2.5e-3, J, ntf)
        eo = O.init_dust_source_fct1(lam, iRT, 3, 2.5e-3, J, tg.xI_scatt, ntf, pola, True)
        assert np.array_equal(eg, eo) and np.abs(eg[:, :, 0]).sum() > 0
        assert (eg[:, :, (4 if pola else 1)] == 0).all()                   # direct stellar light is not part of the source function
        for tau_obs in (1.0e30, 3.0):
            Ig = G.integ_ray_dust(lam, x, y, z, u, v, w, ic, tau_obs, ntf)
            Io = O.integ_ray_dust(lam, x, y, z, u, v, w, ic, tau_obs, eo)
            scale = np.abs(Io).max(axis=1, keepdims=True) + 1e-300
            assert np.abs(Ig - Io).max() / scale.max() < 1e-9
            assert np.mean(np.abs(Ig - Io) <= 1e-11 * scale) > 0.9999          # (a ray on an azimuthal bin edge may flip with atan2's last bit)
            assert (Ig[0] > 0).mean() > 0.9
    with pytest.raises(api.McfostB200Error):
        G.init_dust_source_fct1(lam, 4, 1.0, J, ntf)                      # only 3 observer directions were tallied
    with pytest.raises(api.McfostB200Error):
        G.integ_ray_dust(lam + 1, x, y, z, u, v, w, ic, 1.0e30, ntf)      # source function of another wavelength
    G.close()


@pytest.mark.parametrize("name", ["cyl2D", "cyl3D", "sph2D"])
def test_integ_tau_rays_from_the_star_bit_exact(name):
    """integ_tau (optical_depth.f90:186-244): the two rays it sends from the origin (midplane; the inclination of interest)
    through index_cell + optical_length_tot, identical to the oracle's."""
    P = small_problems()[name]()
    O, G = Oracle(P), api.PhotonLoop(P)
    w0 = np.array([0.0, np.cos(np.float32(60.0) * np.pi / 180.0)]); u0 = np.sqrt(1.0 - w0 * w0); u0[0] = 1.0
    z = np.zeros(2)
    ico, icg = O.index_cell(z, z, z), G.index_cell(z, z, z)
    assert np.array_equal(ico, icg)
    for lam in (1, P.lambda_seuil, P.n_lambda):
        o = O.optical_length_tot(lam, z, z, z, u0, z, w0, ico)
        g = G.optical_length_tot(lam, z, z, z, u0, z, w0, icg)
        assert np.array_equal(o["tau_tot"], g["tau_tot"]) and np.array_equal(o["n_steps"], g["n_steps"]) and np.array_equal(o["lmax"], g["lmax"])
        assert g["tau_tot"][0] > g["tau_tot"][1] > 0 or name == "sph2D"
    G.close()


@pytest.mark.parametrize("name", ["cyl2D", "variable_dust", "cyl3D"])
def test_repartition_energie_on_the_device_matches_oracle(name):
    """mcfost_b200_repartition_energie (thermal_emission.f90:1771-1949, LTE case): E_disk, the star / disk fractions and
    prob_E_cell(0:n_cells, lambda) from a temperature field, against the oracle's sequential sums (parallel scan on the device:
    1e-11; CUDA exp); with a dark zone, with weights, on a wavelength range; then an SED-like step whose disk emission samples
    the device-built table equals the step on uploaded tables."""
    P = S.ref41_multi_like(n_photons_eq_th=100) if name == "variable_dust" else small_problems()[name]()
    rng = np.random.default_rng(4)
    T = rng.uniform(15.0, 900.0, P.n_cells).astype(np.float32); T[::17] = 0.0
    dark = np.zeros(P.n_cells, np.int32); dark[5::11] = 1
    O, G = Oracle(P), api.PhotonLoop(P)
    O.set_dark_zone(dark); G.upload_dark_zone(dark)
    for weight, (l0, l1) in ((None, (1, P.n_lambda)), (rng.uniform(0.5, 2.0, P.n_cells), (7, 19))):
        o = O.repartition_energie(T, P.tab_lambda, P.E_stars, None, weight, l0, l1)
        g = G.repartition_energie(T, P.tab_lambda, P.E_stars, None, weight, l0, l1)
        sl = slice(l0 - 1, l1)
        for k in ("E_disk", "frac_E_stars", "frac_E_disk", "weight_norm"):
            assert np.allclose(g[k][sl], o[k][sl], rtol=1e-11, atol=1e-300), k
        assert np.abs(g["prob_E_cell"][:, sl] - o["prob_E_cell"][:, sl]).max() < 1e-11
        assert (g["prob_E_cell"][0, sl] == 0).all() and np.allclose(g["prob_E_cell"][-1, sl][o["E_disk"][sl] > 0], 1.0, rtol=1e-12)
        assert (np.diff(g["prob_E_cell"][:, sl], axis=0) >= 0).all()
    # the device tables drive the emission: a monochromatic step at a thermal wavelength, disk emission on
    full = G.repartition_energie(T, P.tab_lambda, P.E_stars)
    import copy
    P2 = copy.copy(P); P2.prob_E_cell = None; P2.frac_E_stars = None; P2.frac_E_disk = None
    G.upload_emission(P2)                                # NULL for the three tables: the device-built ones stay
    lam = P.n_lambda - 8
    kw = dict(letape_th=0, lmono=1, lcount_sent=1, n_phot_lim=1.0e30)      # chunks end on packets sent: the same packets in both runs
    p_lam = lam if P.p_n_lambda_pos > 1 else 1
    t_dev = G.mc_photon_loop(lam, p_lam, 300, call_index=5, **kw)
    P3 = copy.copy(P); P3.prob_E_cell = full["prob_E_cell"]; P3.frac_E_stars = full["frac_E_stars"]; P3.frac_E_disk = full["frac_E_disk"]
    G.upload_emission(P3)                                # the same tables through the host
    t_up = G.mc_photon_loop(lam, p_lam, 300, call_index=5, **kw)
    assert t_dev.stats[0] == t_up.stats[0] == 128 * 300 and full["frac_E_stars"][lam - 1] < 0.9      # the disk does emit here
    assert np.array_equal(t_dev.stats[:7], t_up.stats[:7]) and np.allclose(t_dev.sed, t_up.sed, rtol=1e-9, atol=0)
    G.close()
