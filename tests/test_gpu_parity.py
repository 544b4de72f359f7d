"""GPU parity tests (run on the B200 box with -m gpu): every call goes through the
C ABI of libmcfost_b200.so; the oracle is only the checker.

Bars (BASELINE.json north_star): deterministic sub-kernels -- cell indices
bit-exact, lengths within 1e-12 relative (they are in fact required bit-equal
here); Monte Carlo -- temperature median |dT/T| < 1 %, SED within 3 sigma."""
import numpy as np
import pytest

from mcfost_b200 import api, synthetic as S
from oracle.binding import Oracle

from helpers import rays_in_cells, rays_from_outside, small_problems

pytestmark = pytest.mark.gpu

GRIDS = ["cyl2D", "cyl3D", "sph2D", "sph3D"]


@pytest.fixture(scope="module")
def pairs():
    out = {}
    for name in GRIDS:
        P = small_problems()[name]()
        out[name] = (P, Oracle(P), api.PhotonLoop(P))       # upload_grid verifies the cell numbering
    yield out
    for _, _, g in out.values():
        g.close()


@pytest.mark.parametrize("name", GRIDS)
def test_index_cell_bit_exact(pairs, name):
    P, O, G = pairs[name]
    ic, x, y, z, *_ = rays_in_cells(P, 100000, seed=21)
    assert np.array_equal(G.index_cell(x, y, z), O.index_cell(x, y, z))
    # points outside / inside the inner hole / on the axis
    xs, ys, zs, *_ = rays_from_outside(P, 2000)
    x2 = np.concatenate([xs, 0.1 * x[:1000], np.zeros(10)]); y2 = np.concatenate([ys, 0.1 * y[:1000], np.zeros(10)])
    z2 = np.concatenate([zs, 0.1 * z[:1000], np.linspace(-1, 1, 10)])
    assert np.array_equal(G.index_cell(x2, y2, z2), O.index_cell(x2, y2, z2))


@pytest.mark.parametrize("name", GRIDS)
def test_cross_cell_bit_exact(pairs, name):
    P, O, G = pairs[name]
    ic, x, y, z, u, v, w = rays_in_cells(P, 200000, seed=22)
    o, g = O.cross_cell(x, y, z, u, v, w, ic), G.cross_cell(x, y, z, u, v, w, ic)
    assert np.array_equal(g["next_cell"], o["next_cell"])
    for k in ("l", "l_contrib", "l_void_before", "x1", "y1", "z1"):
        assert np.array_equal(g[k], o[k]), k
    # second crossing from the exit points (exercises virtual cells and wall-hugging starts)
    # (cells of the outer radial shell are never crossed: test_exit_grid fires first, optical_depth.f90:87)
    ok = P.cell_map_i[o["next_cell"] - 1] <= P.n_rad
    a = [q[ok] for q in (o["x1"], o["y1"], o["z1"], u, v, w, o["next_cell"], ic)]
    o2, g2 = O.cross_cell(*a), G.cross_cell(*a)
    assert np.array_equal(g2["next_cell"], o2["next_cell"]) and np.array_equal(g2["l"], o2["l"])
    # axis-aligned / degenerate directions
    n = 3000
    uu = np.zeros(n); vv = np.zeros(n); ww = np.zeros(n)
    uu[:1000] = 1.0; vv[1000:2000] = -1.0; ww[2000:] = np.where(np.arange(1000) % 2 == 0, 1.0, -1.0)
    o3 = O.cross_cell(x[:n], y[:n], z[:n], uu, vv, ww, ic[:n]); g3 = G.cross_cell(x[:n], y[:n], z[:n], uu, vv, ww, ic[:n])
    assert np.array_equal(g3["next_cell"], o3["next_cell"]) and np.array_equal(g3["l"], o3["l"])


@pytest.mark.parametrize("name", GRIDS)
def test_move_to_grid_bit_exact(pairs, name):
    P, O, G = pairs[name]
    x, y, z, u, v, w = rays_from_outside(P, 50000)
    o, g = O.move_to_grid(x, y, z, u, v, w), G.move_to_grid(x, y, z, u, v, w)
    assert np.array_equal(g["lintersect"], o["lintersect"]) and np.array_equal(g["icell"], o["icell"])
    for k in ("x", "y", "z"):
        assert np.array_equal(g[k], o[k]), k


@pytest.mark.parametrize("name", GRIDS)
def test_tau_integration_along_fixed_rays(pairs, name):
    P, O, G = pairs[name]
    ic, x, y, z, u, v, w = rays_in_cells(P, 30000, seed=23)
    o = O.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    g = G.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    assert np.array_equal(g["n_steps"], o["n_steps"])                  # identical cell walks
    assert np.allclose(g["tau_tot"], o["tau_tot"], rtol=1e-12, atol=0)
    assert np.allclose(g["lmax"], o["lmax"], rtol=1e-12, atol=0) and np.allclose(g["lmin"], o["lmin"], rtol=1e-12, atol=0)
    tau = np.random.default_rng(3).exponential(3.0, len(x)).astype(np.float32)
    o = O.physical_length(P.lambda_seuil, x, y, z, u, v, w, ic, tau)
    g = G.physical_length(P.lambda_seuil, x, y, z, u, v, w, ic, tau)
    for k in ("flag_sortie", "lpacket_alive", "icell"):
        assert np.array_equal(g[k], o[k]), k
    for k in ("x", "y", "z", "ltot"):
        assert np.allclose(g[k], o[k], rtol=1e-12, atol=0), k


def test_empty_and_error_paths():
    P = small_problems()["cyl2D"]()
    G = api.PhotonLoop(P)
    e = np.zeros(0)
    assert len(G.index_cell(e, e, e)) == 0
    assert len(G.cross_cell(e, e, e, e, e, e, np.zeros(0, np.int32))["l"]) == 0
    with pytest.raises(api.McfostB200Error) as err:
        G.mc_photon_loop(1, 1, 10, lonly_LTE=0)
    assert err.value.code == 6         # MCB_ERR_STATE: nLTE / nRE re-emission needs upload_grains first
    with pytest.raises(api.McfostB200Error) as err:
        G.mc_photon_loop(1, 1, 10, loutput_mc=1, letape_th=0, lmono=1, lmono0=1)
    assert err.value.code == 2         # MCB_ERR_BAD_ARG: photon maps without npix / map_size
    with pytest.raises(api.McfostB200Error) as err:
        G.mc_photon_loop(1, 1, 10, letape_th=0, lmono=1, lscatt_ray_tracing1=1, RT_n_incl=17, RT_n_az=1,
                         tab_u_rt=np.zeros((17, 1)), tab_v_rt=np.zeros((17, 1)), tab_w_rt=np.ones(17))
    assert err.value.code == 5         # MCB_ERR_UNSUPPORTED: fails loudly, no silent fallback
    with pytest.raises(api.McfostB200Error):
        G.mc_photon_loop(P.n_lambda + 1, 1, 10)
    t = G.mc_photon_loop(1, 1, 0)      # zero packets
    assert t.stats[0] == 0 and t.xKJ_abs.sum() == 0
    # a corrupted cell map is rejected
    P2 = small_problems()["cyl2D"]()
    P2.cell_map_i = P2.cell_map_i.copy(); P2.cell_map_i[3] += 1
    with pytest.raises(api.McfostB200Error) as err:
        api.PhotonLoop(P2)
    assert err.value.code == 3
    G.close()


def test_dark_zone_definition_matches_oracle():
    """define_dark_zone's ray walk (optical_depth.f90:1519-1550) on the GPU vs the oracle."""
    P = S.ref41_like(n_photons_eq_th=100, dark_zone=False)
    O, G = Oracle(P), api.PhotonLoop(P)
    do = S.define_dark_zone(P, P.lambda_seuil, 1500.0, O.dark_zone_walker())
    dg = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
    assert do.sum() > 0 and np.array_equal(do, dg)
    G.close()


def _thermal_pair(P, n2):
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False)
    G.close()
    to = Oracle(P, fast=True).run(n_threads=0, n_photons2=n2)
    return to, tg


def test_thermal_ref41_like_statistical_parity():
    """G1 at the file-default budget (128 x 1000 packets): temperature and escaping SED."""
    P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
    P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, Oracle(P).dark_zone_walker())
    S.repartition_energie(P)
    to, tg = _thermal_pair(P, 1000)
    assert tg.stats[0] == to.stats[0] == 128000 == tg.n_phot_envoyes.sum()
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]                       # every packet detected or killed
    assert tg.sed.sum() == pytest.approx(tg.stats[6])                     # energy conservation (E_paquet = 1)
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0) & (P.l_dark_zone == 0)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert lit.sum() > 0.9 * (P.l_dark_zone == 0).sum()
    assert np.median(rel) < 0.01, np.median(rel)                          # north_star bar
    assert np.percentile(rel, 75) < 0.05                                   # the reference's own MC_similar bar (test_mcfost.py:88)
    # SED per (lambda, inclination) bin within 3 sigma of packet statistics (Poisson on counts, two samples)
    no, ng = to.n_phot_sed[:, :, 0], tg.n_phot_sed[:, :, 0]
    m = (no + ng) > 50
    zscore = (ng[m] - no[m]) / np.sqrt(no[m] + ng[m])
    assert np.mean(np.abs(zscore) < 3) > 0.99 and abs(zscore.mean()) < 0.3
    # global absorbed energy
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.01


@pytest.mark.parametrize("name", ["cyl3D", "sph2D", "sph3D"])
def test_thermal_other_grids_statistical_parity(name, monkeypatch):
    # (production launch configuration: the library itself keeps the packets in flight a small fraction of the packets
    # sent)
    P = small_problems()[name]()
    to, tg = _thermal_pair(P, 1500)
    assert tg.stats[0] == to.stats[0]
    assert tg.sed.sum() == pytest.approx(tg.stats[6])
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.03            # two independent MC estimates (192k packets each)
    assert abs(tg.stats[1] / to.stats[1] - 1) < 0.02                      # cell-crossing steps
    no, ng = to.n_phot_sed.sum(axis=(0, 2)), tg.n_phot_sed.sum(axis=(0, 2))
    assert (np.abs(ng - no) < 4 * np.sqrt(no + ng) + 1).all()


def _sed_kwargs(pola):
    return dict(letape_th=0, lmono=1, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=200.0,
                lscatt_ray_tracing1=1, lsepar_pola=int(pola), lsepar_contrib=1, RT_n_incl=3, RT_n_az=1,
                tab_u_rt=np.array([[0.0], [0.5], [0.9]]), tab_v_rt=np.zeros((3, 1)),
                tab_w_rt=np.array([1.0, np.sqrt(0.75), np.sqrt(0.19)]))


@pytest.mark.parametrize("pola", [False, True])
def test_sed_step_matches_oracle_packet_by_packet(pola):
    """Forced-scattering (lmono) mode has no feedback: with the shared per-packet Philox streams the
    GPU follows the same trajectories as the oracle; tallies agree up to libm / summation-order noise."""
    P = small_problems()["cyl2D"]()
    kw = _sed_kwargs(pola)
    ntf = (4 if pola else 1) + 4
    n_xI = 45 * 2 * ntf * 3 * P.n_cells
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(kw.pop("lambda_in"), kw.pop("p_lambda_in"), kw.pop("n_photons2"), kw.pop("n_phot_lim"), 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, n_xI=n_xI, **_sed_kwargs(pola))
    assert tg.stats[0] == to.stats[0] == 128 * 200
    assert abs(tg.stats[1] - to.stats[1]) <= 1e-4 * to.stats[1]           # same walks (a few libm-ulp flips allowed)
    assert abs(tg.stats[2] - to.stats[2]) <= 1e-4 * to.stats[2]
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=2e-3)
    if pola:
        assert np.abs(tg.sed_q).sum() > 0
        assert np.allclose(tg.sed_q.sum(axis=(0, 2)), to.sed_q.sum(axis=(0, 2)), rtol=0.02, atol=1e-3 * np.abs(to.sed_q).sum())
    xg = tg.xI_scatt.reshape((45, 2, ntf, 3, P.n_cells), order="F")
    xo = to.xI_scatt.reshape((45, 2, ntf, 3, P.n_cells), order="F")
    assert xo[:, :, 0].sum() > 0
    # per observer direction and per flux type, summed over cells: fp32 accumulation noise only
    assert np.allclose(xg.sum(axis=(0, 1, 4)), xo.sum(axis=(0, 1, 4)), rtol=5e-3, atol=1e-6 * np.abs(xo).sum())
    # cell-resolved for the bright cells
    cg, co = xg[:, :, 0].sum(axis=(0, 1, 2)), xo[:, :, 0].sum(axis=(0, 1, 2))
    bright = co > 0.01 * co.max()
    assert np.allclose(cg[bright], co[bright], rtol=0.02)


def test_sed_chunk_termination_counts_received_packets():
    """SED mode stops a chunk once n_photons2 packets were received in bin capt_sup
    (dust_transfer.f90:510,529,551); packets in flight finish, so >= the target."""
    P = small_problems()["cyl2D"]()
    G = api.PhotonLoop(P)
    t = G.mc_photon_loop(8, 8, 5, 1.0e4, 1, False, letape_th=0, lmono=1)
    G.close()
    recv_bin2 = t.n_phot_sed[7, 1, 0]
    assert recv_bin2 >= 128 * 5
    assert t.stats[0] == t.n_phot_envoyes[7] and t.stats[0] < 128 * 1.0e4


def test_rank_partition_is_additive():
    """Two 'ranks' on one device: chunk round-robin + identical Philox streams => tallies add up."""
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1)
    G = api.PhotonLoop(P)
    full = G.mc_photon_loop(8, 8, 10 ** 9, 64.0, 1, False, **kw)
    a = G.mc_photon_loop(8, 8, 10 ** 9, 64.0, 1, False, rank=0, n_ranks=2, **kw)
    b = G.mc_photon_loop(8, 8, 10 ** 9, 64.0, 1, False, rank=1, n_ranks=2, **kw)
    G.close()
    assert a.stats[0] + b.stats[0] == full.stats[0] == 128 * 64
    assert np.array_equal(a.n_phot_sed + b.n_phot_sed, full.n_phot_sed)
    assert np.allclose(a.sed + b.sed, full.sed, rtol=1e-9, atol=1e-12)


def test_full_size_properties_ref41():
    """BASELINE-size run (G1, 128 x 20000 packets): size-independent properties."""
    P = S.ref41_like(n_photons_eq_th=20000, dark_zone=True)       # steps 1-3 dark zone (superset), fine for properties
    G = api.PhotonLoop(P)
    t = G.mc_photon_loop(1, 1, 20000, 1.0e30, 1, False)
    G.close()
    assert t.stats[0] == 128 * 20000 == t.n_phot_envoyes.sum()
    assert t.stats[5] + t.stats[6] == t.stats[0]
    assert t.sed.sum() == pytest.approx(t.stats[6], rel=1e-9)
    assert (t.sed_star + t.sed_star_scat + t.sed_disk + t.sed_disk_scat).sum() == pytest.approx(t.sed.sum(), rel=1e-9)
    assert (t.xKJ_abs[P.l_dark_zone == 1] == 0).all()              # nothing is deposited inside the dark zone
    assert (t.xT_ech >= 2).all() and (t.xT_ech <= P.n_T).all()
    T = S.temp_finale(P, t.xKJ_abs)
    assert 100 < T.max() < 3000
    # radiative equilibrium: total re-emitted = absorbed => detected energy equals emitted energy
    assert t.sed.sum() / t.stats[0] > 0.999


@pytest.fixture(scope="module")
def voronoi_pair():
    P = S.voronoi_disk(n_points=1500, n_photons_eq_th=300)
    G = api.PhotonLoop(P)
    yield P, Oracle(P), G
    G.close()


def _voronoi_rays(P, n, seed):
    rng = np.random.default_rng(seed)
    ic = rng.integers(1, P.n_cells + 1, n).astype(np.int32)
    x, y, z = P.vor_xyz[0, ic - 1].copy(), P.vor_xyz[1, ic - 1].copy(), P.vor_xyz[2, ic - 1].copy()
    # jitter inside the cell: move 30 % of the way towards a random direction's wall
    w = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2 * np.pi, n)
    u, v = np.sqrt(1 - w * w) * np.cos(ph), np.sqrt(1 - w * w) * np.sin(ph)
    return ic, x, y, z, u, v, w


def test_voronoi_deterministic_kernels_bit_exact(voronoi_pair):
    """fp32 plane tests, wall planes and the cut-cell sphere of cross_Voronoi_cell (Voronoi.f90:839-992)."""
    P, O, G = voronoi_pair
    ic, x, y, z, u, v, w = _voronoi_rays(P, 100000, 31)
    o, g = O.cross_cell(x, y, z, u, v, w, ic), G.cross_cell(x, y, z, u, v, w, ic)
    assert np.array_equal(g["next_cell"], o["next_cell"])
    for k in ("l", "l_contrib", "l_void_before", "x1", "y1", "z1"):
        assert np.array_equal(g[k], o[k]), k
    # continue from the exit points with previous_cell set (the entry face is skipped, :874)
    ok = o["next_cell"] > 0
    a = [q[ok] for q in (o["x1"], o["y1"], o["z1"], u, v, w, o["next_cell"], ic)]
    o2, g2 = O.cross_cell(*a), G.cross_cell(*a)
    assert np.array_equal(g2["next_cell"], o2["next_cell"]) and np.array_equal(g2["l"], o2["l"])
    assert np.array_equal(G.index_cell(x[:2000], y[:2000], z[:2000]), O.index_cell(x[:2000], y[:2000], z[:2000]))
    o = O.optical_length_tot(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000])
    g = G.optical_length_tot(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000])
    assert np.array_equal(g["n_steps"], o["n_steps"])
    assert np.allclose(g["tau_tot"], o["tau_tot"], rtol=1e-12, atol=0) and np.allclose(g["lmax"], o["lmax"], rtol=1e-12, atol=0)
    tau = np.random.default_rng(4).exponential(2.0, 20000).astype(np.float32)
    o = O.physical_length(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000], tau)
    g = G.physical_length(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000], tau)
    for k in ("flag_sortie", "lpacket_alive", "icell"):
        assert np.array_equal(g[k], o[k]), k
    assert np.allclose(g["x"], o["x"], rtol=1e-12, atol=0)
    # entry from outside the box
    xs, ys, zs, du, dv, dw = rays_from_outside(P, 3000)
    o, g = O.move_to_grid(xs, ys, zs, du, dv, dw), G.move_to_grid(xs, ys, zs, du, dv, dw)
    assert np.array_equal(g["lintersect"], o["lintersect"]) and np.array_equal(g["icell"], o["icell"])


def test_voronoi_thermal_statistical_parity(voronoi_pair, monkeypatch):
    P, O, G = voronoi_pair
    tg = G.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False)
    to = Oracle(P, fast=True).run(n_threads=0, n_photons2=1500)
    assert tg.stats[0] == to.stats[0] == 128 * 1500
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]
    assert tg.sed.sum() == pytest.approx(tg.stats[6])
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.03
    assert abs(tg.stats[1] / to.stats[1] - 1) < 0.02
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
    assert np.median(np.abs(Tg[lit] - To[lit]) / To[lit]) < 0.02          # 1500 cells, 192k packets: MC noise dominated


def test_variable_dust_per_cell_tables(monkeypatch):
    """lvariable_dust = .true. (ref4.1_multi-like, LTE part): every opacity / scattering / thermal table is
    indexed by cell (p_n_cells = n_cells, kappa_factor = 1), single-wavelength scattering tables
    (p_n_lambda_pos = 1); the kernel reads them from global memory instead of the shared-memory staging."""
    P = S.ref41_multi_like(n_photons_eq_th=1500)
    O, G = Oracle(P), api.PhotonLoop(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 20000, seed=41)
    o = O.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    g = G.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    assert np.array_equal(g["n_steps"], o["n_steps"]) and np.allclose(g["tau_tot"], o["tau_tot"], rtol=1e-12, atol=0)
    tg = G.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False)
    G.close()
    to = Oracle(P, fast=True).run(n_threads=0, n_photons2=1500)
    assert tg.stats[0] == to.stats[0] == 128 * 1500
    assert tg.sed.sum() == pytest.approx(tg.stats[6])
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.03
    assert abs(tg.stats[3] / to.stats[3] - 1) < 0.05 and abs(tg.stats[4] / to.stats[4] - 1) < 0.05   # scatterings, absorptions
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
    assert np.median(np.abs(Tg[lit] - To[lit]) / To[lit]) < 0.02


def test_full_size_3d_grid_properties():
    """ref4.1_3D-size grid: 100 x (2 x 50) x 72 = 720 000 cells, no dark zone (dust_transfer.f90:290-293)."""
    P = S.ref41_3d_like(n_photons_eq_th=2000, tau_mid=300.0)
    assert P.n_cells == 720000
    G = api.PhotonLoop(P)                      # upload verifies the closed-form numbering of all 734 k ids
    ic, x, y, z, u, v, w = rays_in_cells(P, 50000, seed=51)
    assert np.array_equal(G.index_cell(x, y, z), ic)
    t = G.mc_photon_loop(1, 1, 2000, 1.0e30, 1, False)
    G.close()
    assert t.stats[0] == 128 * 2000 == t.n_phot_envoyes.sum()
    assert t.stats[5] + t.stats[6] == t.stats[0]
    assert t.sed.sum() == pytest.approx(t.stats[6], rel=1e-9)
    # azimuthal structure of the m=2 spiral shows up in the absorbed energy, top / bottom halves agree statistically
    e = t.xKJ_abs.reshape((72, 100, 100))          # (k, j-row, i)
    top, bot = e[:, 50:, :].sum(), e[:, :50, :].sum()
    assert abs(top / bot - 1) < 0.05


@pytest.mark.parametrize("pola", [False, True])
def test_image_step_rt2_matches_oracle(pola):
    """run_image_mc's call (dust_transfer.f90:758): lmono0, fixed packet count, rt2 accumulators I_spec /
    I_spec_star (radiation_field.f90:91-130); no feedback, so trajectories match the oracle packet by packet."""
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, lmono0=1, lscatt_ray_tracing2=1, lsepar_pola=int(pola), lsepar_contrib=1)
    ntf = (4 if pola else 1) + 4
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(8, 8, 300, 1.0e30, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, n_Ispec=ntf * 15 * 15 * P.n_cells, lambda_in=8, p_lambda_in=8, n_photons2=300, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 300
    assert tg.sed.sum() == 0 and to.sed.sum() == 0                        # image mode keeps no SED (output.f90:360)
    assert abs(tg.stats[1] - to.stats[1]) <= 1e-4 * to.stats[1]
    Ig = tg.I_spec.reshape((ntf, 15, 15, P.n_cells), order="F"); Io = to.I_spec.reshape((ntf, 15, 15, P.n_cells), order="F")
    assert Io[0].sum() > 0 and to.I_spec_star.sum() > 0
    assert np.allclose(tg.I_spec_star.sum(), to.I_spec_star.sum(), rtol=2e-3)
    assert np.allclose(Ig.sum(axis=(1, 2, 3)), Io.sum(axis=(1, 2, 3)), rtol=5e-3, atol=1e-6 * np.abs(Io).sum())
    assert np.allclose(Ig[0].sum(axis=(1, 2)), Io[0].sum(axis=(1, 2)), rtol=0.01, atol=1e-4 * Io[0].sum())   # (theta_I) profile
    bright = Io[0].sum(axis=(0, 1)) > 0.01 * Io[0].sum(axis=(0, 1)).max()
    assert np.allclose(Ig[0].sum(axis=(0, 1))[bright], Io[0].sum(axis=(0, 1))[bright], rtol=0.02)


def test_straggler_handover_keeps_every_packet():
    """A thermal call above the small-budget threshold runs on three launches: the packet-per-warp kernel sends the first
    packets, the packet-per-lane kernel the bulk, and it parks its last packets for the packet-per-warp kernel.  Nothing
    is lost or duplicated, with and without mcfost_b200_set_overlap, and the tallies are those of a small-budget call
    (packet-per-warp kernel alone) within Monte Carlo noise."""
    P = small_problems()["cyl2D"]()
    G = api.PhotonLoop(P)
    n2 = 12000
    small = G.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False, lsepar_pola=1)
    assert G.debug_counters()["parked"] == 0      # 192 000 packets: the low-latency kernel alone
    ref = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False, lsepar_pola=1)
    d0 = G.debug_counters()
    G.set_overlap(8)
    t = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False, lsepar_pola=1, call_index=1)
    d = G.debug_counters()
    for q in (ref, t):
        assert q.stats[0] == 128 * n2 == q.n_phot_envoyes.sum()
        assert q.stats[5] + q.stats[6] == q.stats[0]
        assert q.sed.sum() == pytest.approx(q.stats[6])
    assert abs(t.xKJ_abs.sum() / ref.xKJ_abs.sum() - 1) < 0.01
    assert abs(t.stats[1] / ref.stats[1] - 1) < 0.01 and abs(t.stats[2] / ref.stats[2] - 1) < 0.01
    assert np.abs(t.sed_q).sum() > 0 and abs(np.abs(t.sed_q).sum() / np.abs(ref.sed_q).sum() - 1) < 0.1
    assert d["parked"] > 0 and d0["parked"] > 0   # the hand-over did happen
    # per packet the three-launch call and the one-launch call are the same physics
    # (the two budgets see different running temperatures, hence slightly different wavelengths and path lengths)
    assert abs(ref.stats[1] / ref.stats[0] / (small.stats[1] / small.stats[0]) - 1) < 0.06
    assert abs(ref.xKJ_abs.sum() / ref.stats[0] / (small.xKJ_abs.sum() / small.stats[0]) - 1) < 0.06
    # SED mode counts received packets: no hand-over there, the call is unchanged
    s = G.mc_photon_loop(8, 8, 10 ** 9, 64.0, 1, False, letape_th=0, lmono=1)
    assert s.stats[0] == 128 * 64
    with pytest.raises(api.McfostB200Error):
        G.set_overlap(-1)
    G.close()


def test_more_handles_than_constant_banks_run_concurrently():
    """Four handles launch back to back (the library has three constant banks): the bank guard serialises the
    two that share a bank, every call keeps its own tallies and conserves its packets."""
    P = small_problems()["cyl2D"]()
    Gs = [api.PhotonLoop(P) for _ in range(4)]
    for g in Gs:
        g.set_overlap(16, 8)
    runs = [g.launch(1, 1, 3000 + 500 * i, 1.0e30, 1, call_index=i) for i, g in enumerate(Gs)]
    for i, (g, r) in enumerate(zip(Gs, runs)):
        g.sync()
        t = g.download(r)
        assert t.stats[0] == 128 * (3000 + 500 * i) == t.n_phot_envoyes.sum()
        assert t.stats[5] + t.stats[6] == t.stats[0]
        assert t.sed.sum() == pytest.approx(t.stats[6])
    # same seed, same call_index on two different handles: identical packets, tallies equal up to summation order
    a = Gs[0].mc_photon_loop(8, 8, 10 ** 9, 32.0, 1, False, letape_th=0, lmono=1, call_index=7)
    b = Gs[3].mc_photon_loop(8, 8, 10 ** 9, 32.0, 1, False, letape_th=0, lmono=1, call_index=7)
    assert np.array_equal(a.n_phot_sed, b.n_phot_sed) and np.allclose(a.sed, b.sed, rtol=1e-9, atol=1e-12)
    for g in Gs:
        g.close()


@pytest.mark.parametrize("sym", [0, 1])
def test_photon_maps_match_oracle(sym):
    """Image step with loutput_mc: forced scattering has no feedback, so with the shared Philox streams every packet
    lands in the same pixel as in the oracle (up to libm-ulp flips at pixel edges)."""
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, lmono0=1, loutput_mc=1, lsepar_pola=1, lsepar_contrib=1, l_sym_ima=sym,
              npix_x=24, npix_y=16, map_size=700.0, N_thet=4, N_phi=2)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(8, 8, 400, 1.0e30, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, lambda_in=8, p_lambda_in=8, n_photons2=400, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 400
    mg, mo = tg.stokes_map, to.stokes_map
    assert mg.shape == mo.shape == (24, 16, 4, 2, 8)
    assert mo[..., 0].sum() > 0 and np.isclose(mg[..., 0].sum(), mo[..., 0].sum(), rtol=1e-3)
    # per detector bin and per plane
    assert np.allclose(mg.sum(axis=(0, 1)), mo.sum(axis=(0, 1)), rtol=5e-3, atol=2e-3 * np.abs(mo).sum(axis=(0, 1)).max())
    # pixel by pixel for the bright pixels of the intensity map
    Ig, Io = mg[..., 0], mo[..., 0]
    bright = Io > 0.01 * Io.max()
    assert bright.sum() > 20
    assert np.allclose(Ig[bright], Io[bright], rtol=0.02)
    # the four contributions still add up to I on the device
    assert np.allclose(mg[..., 4:8].sum(axis=-1), mg[..., 0], rtol=1e-9, atol=1e-12)


def test_capt_interet_and_origin_match_oracle():
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, N_thet=5, lorigine=1, capt_interet=2, lonly_capt_interet=1, capt_inf=2, capt_sup=3)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(8, 8, 10 ** 9, 150.0, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=150.0, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 150
    assert tg.sed[:, 0].sum() == 0 and tg.sed[:, 3:].sum() == 0
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert np.isclose(tg.star_origin[7], to.star_origin[7], rtol=1e-3) and to.star_origin[7] > 0
    assert np.isclose(tg.disk_origin[7].sum(), to.disk_origin[7].sum(), rtol=5e-3)
    assert np.isclose(tg.star_origin[7] + tg.disk_origin[7].sum(), tg.sed[7, 1].sum(), rtol=1e-9)
    big = to.disk_origin[7] > 0.01 * to.disk_origin[7].max()
    assert np.allclose(tg.disk_origin[7][big], to.disk_origin[7][big], rtol=0.05)


def test_packet_counts_per_cell_match_oracle():
    P = small_problems()["cyl2D"]()
    kw = dict(letape_th=0, lmono=1, lxJ_abs=1, lxN_abs=1)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(8, 8, 10 ** 9, 200.0, 1, False, **kw)
    th = G.mc_photon_loop(1, 1, 200, 1.0e30, 1, False, lxN_abs=1)
    G.close()
    to = Oracle(P).run(n_threads=0, xJ=True, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=200.0, **kw)
    assert tg.xN_abs.shape == to.xN_abs.shape == (P.n_cells, P.n_lambda)
    assert to.xN_abs.sum() > 0 and abs(tg.xN_abs.sum() - to.xN_abs.sum()) <= 1e-4 * to.xN_abs.sum()
    assert np.allclose(tg.xN_abs[:, 7], to.xN_abs[:, 7], atol=3, rtol=1e-3)
    assert th.xN_abs.shape == (P.n_cells, 1) and 0 < th.xN_abs.sum() <= th.stats[1]
    assert np.array_equal(th.xN_abs[:, 0] > 0, th.xKJ_abs > 0)


def test_interstellar_radiation_field_matches_oracle():
    """emit_packet_ISM + move_to_grid from outside, packet by packet in the forced-scattering step; ISM packets are
    never detected (flag_ISM) but feed xJ_abs."""
    P = small_problems()["cyl2D"]()
    P.E_ISM = 0.5 * P.E_stars
    P.R_ISM = 1.5 * float(np.sqrt(P.Rmax2)); P.centre_ISM = (0.0, 0.0, 0.0)
    S.repartition_energie(P)
    kw = dict(letape_th=0, lmono=1, lxJ_abs=1)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(8, 8, 10 ** 9, 200.0, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, xJ=True, lambda_in=8, p_lambda_in=8, n_photons2=10 ** 9, n_phot_lim=200.0, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 200
    n_ism = to.stats[0] - to.stats[5] - to.stats[6]
    assert n_ism > 0.2 * to.stats[0]                                        # a third of the packets come from outside
    assert abs(tg.stats[6] - to.stats[6]) <= 3 and abs(tg.stats[5] - to.stats[5]) <= 3
    assert abs(tg.stats[1] - to.stats[1]) <= 2e-4 * to.stats[1]
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=2e-3)
    assert np.isclose(tg.xJ_abs.sum(), to.xJ_abs.sum(), rtol=1e-3)
