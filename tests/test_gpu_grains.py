"""GPU parity of the per-grain branches (run on the B200 box with -m gpu), all through the C ABI:
scattering method 1 (dust_transfer.f90:1291-1317) and the nLTE / qRE re-emission branches
(dust_transfer.f90:1353-1395).  The oracle is only the checker."""
import numpy as np
import pytest

from mcfost_b200 import api, synthetic as S
from oracle.binding import Oracle

pytestmark = pytest.mark.gpu

MIXED = dict(lonly_LTE=0, lRE_nLTE=1, lnRE=1, lxJ_abs_step1=1)


@pytest.mark.parametrize("low_mem", [1, 0])
@pytest.mark.parametrize("aniso1", [1, 0])
def test_method1_forced_scattering_packet_by_packet(low_mem, aniso1):
    """lmono (no feedback): same Philox words -> same grains, same angles, same walks as the oracle."""
    P = S.multi_grain_like(n_photons_eq_th=50, tau_mid=300.0)
    kw = dict(letape_th=0, lmono=1, lscattering_method1=1, lmethod_aniso1=aniso1, lsepar_pola=aniso1,
              low_mem_scattering=low_mem)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(6, 6, 10 ** 9, 300.0, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, lambda_in=6, p_lambda_in=6, n_photons2=10 ** 9, n_phot_lim=300.0, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 300
    assert to.stats[3] > to.stats[0]                                        # more than one scattering per packet on average
    assert abs(tg.stats[1] - to.stats[1]) <= 2e-4 * to.stats[1]
    assert abs(tg.stats[3] - to.stats[3]) <= 2e-4 * to.stats[3]
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=2e-3)
    if aniso1:
        assert np.abs(to.sed_q).sum() > 0
        assert np.allclose(tg.sed_q.sum(axis=(0, 2)), to.sed_q.sum(axis=(0, 2)), rtol=0.02, atol=2e-3 * np.abs(to.sed_q).sum())
        assert np.allclose(tg.sed_u.sum(axis=(0, 2)), to.sed_u.sum(axis=(0, 2)), rtol=0.02, atol=2e-3 * np.abs(to.sed_q).sum())


def test_method1_nontrivial_s11_scales_the_packet_energy():
    """get_Mueller_matrix_per_grain feeds M(1,1) into update_Stokes' energy normalisation
    (scattering.f90:1294): with tab_s11 != 1 the packet energy changes at every scattering."""
    P = S.multi_grain_like(n_photons_eq_th=50, tau_mid=5.0)
    theta = np.arange(181) * np.pi / 180
    P.tab_s11 = np.asfortranarray(np.float32(P.tab_s11 * (1.0 + 0.05 * np.cos(theta))[:, None, None]))
    kw = dict(letape_th=0, lmono=1, lscattering_method1=1, lmethod_aniso1=1, lsepar_pola=1)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(6, 6, 10 ** 9, 200.0, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, lambda_in=6, p_lambda_in=6, n_photons2=10 ** 9, n_phot_lim=200.0, **kw)
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=2e-3)


def test_method1_thermal_statistical_parity(monkeypatch):
    P = S.multi_grain_like(n_photons_eq_th=1500, tau_mid=30.0)
    kw = dict(n_photons2=1500, lscattering_method1=1, lmethod_aniso1=1, lsepar_pola=1, lonly_LTE=1)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, kw.pop("n_photons2"), 1.0e30, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, n_photons2=1500, **kw)
    assert tg.stats[0] == to.stats[0] and tg.stats[5] + tg.stats[6] == tg.stats[0]
    assert _close(tg.stats[3], to.stats[3]) and _close(tg.stats[4], to.stats[4])
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05


def _grain_temperatures(P, t, regime):
    """Final per-grain temperatures from xJ_abs (what temp_finale_nLTE does, thermal_emission.f90:1010-1100)."""
    if regime == "nLTE":
        ks = np.arange(P.grain_RE_nLTE_start, P.grain_RE_nLTE_end + 1); logE = P.log_E_em_1grain
    else:
        ks = np.arange(P.grain_nRE_start, P.grain_nRE_end + 1); logE = P.log_E_em_1grain_nRE
    C = P.C_abs_norm[ks - 1].astype(np.float64)                               # (nk, nl)
    E = (C @ (t.xJ_abs + P.J0).T) * P.L_packet_th / P.volume[None, :]          # (nk, n_cells)
    lt = np.log(P.tab_Temp.astype(np.float64))
    return np.exp(np.stack([np.interp(np.log(E[j]), logE[j], lt) for j in range(len(ks))]))


def _close(a, b, rel=0.01, nsig=4.0):
    """Two independent Monte Carlo counts agree: Poisson noise + a small relative slack."""
    return abs(a - b) <= nsig * np.sqrt(a + b) + rel * b


@pytest.mark.parametrize("variable,low_mem", [(False, 0), (False, 1), (True, 1)])
def test_mixed_heating_regimes_statistical_parity(variable, low_mem, monkeypatch):
    """LTE + nLTE + qRE grains in one thermal step: the three re-emission branches, E_abs_nRE, xJ_abs and
    the per-grain temperature indices against an independent oracle run of the same size.

    Immediate re-emission reads the RUNNING tallies (Bjorkman & Wood): its spectrum is only independent of
    the order of the packets when few of them are in flight at once compared with the packet budget.  The
    full grid keeps 148 x 1024 packets in flight -- nothing against the 1e7-1e9 packets of a production
    run, but most of this test's 192 000: the library picks the number of blocks from the budget."""
    P = S.multi_grain_like(n_photons_eq_th=1500, tau_mid=30.0, variable=variable, pola=False)
    kw = dict(low_mem_th_emission_nLTE=low_mem, **MIXED)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, xJ=True, n_photons2=1500, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 1500
    assert tg.stats[5] + tg.stats[6] == tg.stats[0]
    for a in (1, 3, 4):                                                       # steps, scatterings, absorptions
        assert _close(tg.stats[a], to.stats[a]), (a, tg.stats[a], to.stats[a])
    assert to.E_abs_nRE[0] > 0 and abs(tg.E_abs_nRE[0] / to.E_abs_nRE[0] - 1) < 0.02
    assert abs(tg.xKJ_abs.sum() / to.xKJ_abs.sum() - 1) < 0.02
    assert abs(tg.xJ_abs.sum() / to.xJ_abs.sum() - 1) < 0.02
    # LTE grains
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05
    # nLTE and qRE grains: temperatures from the final xJ_abs
    for regime in ("nLTE", "nRE"):
        a, b = _grain_temperatures(P, to, regime), _grain_temperatures(P, tg, regime)
        rel = np.abs(b - a) / a
        assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05, regime
    # running temperature indices: same shape, same range, same population of touched (grain, cell) pairs
    for name in ("xT_ech_1grain", "xT_ech_1grain_nRE"):
        xo, xg = getattr(to, name), getattr(tg, name)
        assert xg.shape == xo.shape and xg.min() >= 2 and xg.max() <= P.n_T
        both = (xo > 2) & (xg > 2)
        assert both.sum() > 0.8 * (xo > 2).sum()
        assert np.median(np.abs(xg[both] - xo[both])) <= 1
    # qRE mask respected: grains not at equilibrium in a cell never re-emit there
    assert (tg.xT_ech_1grain_nRE[P.l_RE == 0] == 2).all() and (to.xT_ech_1grain_nRE[P.l_RE == 0] == 2).all()
    # emergent spectrum: wavelength bins of the escaping packets (re-emission wavelength sampling)
    no, ng = to.n_phot_sed.sum(axis=(1, 2)), tg.n_phot_sed.sum(axis=(1, 2))
    m = (no + ng) > 100
    z = (ng[m] - no[m]) / np.sqrt(no[m] + ng[m])
    assert np.mean(np.abs(z) < 3.5) > 0.95 and abs(z.mean()) < 0.5


def test_only_nlte_and_state_errors(monkeypatch):
    P = S.multi_grain_like(n_photons_eq_th=400, tau_mid=5.0, pola=False)
    G = api.PhotonLoop(P)
    kw = dict(lonly_LTE=0, lonly_nLTE=1, lRE_nLTE=1, lxJ_abs_step1=1)
    tg = G.mc_photon_loop(1, 1, 400, 1.0e30, 1, False, **kw)
    to = Oracle(P).run(n_threads=0, xJ=True, n_photons2=400, **kw)
    assert _close(tg.stats[4], to.stats[4])
    assert (tg.xT_ech_1grain[0] == 2).all()               # bisection quirk: the first nLTE grain is never chosen
    a, b = _grain_temperatures(P, to, "nLTE"), _grain_temperatures(P, tg, "nLTE")
    assert np.median(np.abs(b - a) / a) < 0.02
    # nLTE re-emission reads xJ_abs: refusing to run without it is a loud error, not a silent zero
    with pytest.raises(api.McfostB200Error) as ei:
        G.mc_photon_loop(1, 1, 10, 1.0e30, 1, False, lonly_LTE=0, lonly_nLTE=1, lRE_nLTE=1)
    assert ei.value.code == 2
    G.close()
    # per-grain modes without upload_grains
    P2 = S.multi_grain_like(n_photons_eq_th=10)
    del P2.n_grains_tot
    G2 = api.PhotonLoop(P2)
    with pytest.raises(api.McfostB200Error) as ei:
        G2.mc_photon_loop(1, 1, 10, 1.0e30, 1, False, lscattering_method1=1)
    assert ei.value.code == 6
    G2.close()


def test_hot_spot_and_weighted_emission_match_oracle_packet_by_packet():
    P = S.multi_grain_like(n_photons_eq_th=50, tau_mid=5.0, pola=False)
    S.repartition_energie(P, Tdust=np.full(P.n_cells, 150.0))        # warm disk: some packets come from the dust
    P.correct_E_emission = 0.25 + 0.5 * np.random.default_rng(3).random(P.n_cells)
    lam = int(np.argmax(P.tab_lambda > 30.0)) + 1
    kw = dict(letape_th=0, lmono=1, lspot=1, T_spot=12000.0, surf_fraction_spot=0.75, theta_spot=30.0, phi_spot=40.0,
              star1_T=float(P.star_T[0]), tab_lambda=P.tab_lambda, lweight_emission=1, l_sym_axiale=0, N_phi=4)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(lam, lam, 10 ** 9, 300.0, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, lambda_in=lam, p_lambda_in=lam, n_photons2=10 ** 9, n_phot_lim=300.0, **kw)
    assert tg.stats[0] == to.stats[0] == 128 * 300
    assert np.allclose(tg.n_phot_sed, to.n_phot_sed, atol=3)
    assert to.sed_disk.sum() > 0 and to.sed_star.sum() > 0
    assert np.isclose(tg.sed_star.sum(), to.sed_star.sum(), rtol=2e-3)
    assert np.isclose(tg.sed_disk.sum(), to.sed_disk.sum(), rtol=2e-3)
    assert np.allclose(tg.sed.sum(axis=0), to.sed.sum(axis=0), rtol=5e-3, atol=1e-3 * to.sed.sum())


def test_low_memory_lte_emission_statistical_parity(monkeypatch):
    P = S.multi_grain_like(n_photons_eq_th=1500, tau_mid=30.0, pola=False)
    kw = dict(lonly_LTE=1, low_mem_th_emission=1)
    G = api.PhotonLoop(P)
    tg = G.mc_photon_loop(1, 1, 1500, 1.0e30, 1, False, **kw)
    G.close()
    to = Oracle(P).run(n_threads=0, n_photons2=1500, **kw)
    assert tg.stats[0] == to.stats[0] and tg.stats[5] + tg.stats[6] == tg.stats[0]
    assert _close(tg.stats[4], to.stats[4]) and _close(tg.stats[3], to.stats[3])
    To, Tg = S.temp_finale(P, to.xKJ_abs), S.temp_finale(P, tg.xKJ_abs)
    lit = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
    rel = np.abs(Tg[lit] - To[lit]) / To[lit]
    assert np.median(rel) < 0.01 and np.percentile(rel, 75) < 0.05
    no, ng = to.n_phot_sed.sum(axis=(1, 2)), tg.n_phot_sed.sum(axis=(1, 2))
    m = (no + ng) > 100
    z = (ng[m] - no[m]) / np.sqrt(no[m] + ng[m])
    assert np.mean(np.abs(z) < 3.5) > 0.95 and abs(z.mean()) < 0.5


def test_temp_finale_on_device_matches_the_tallies(monkeypatch):
    """mcfost_b200_temp_finale / _temp_finale_nlte read the device-resident tallies of the last call: they must be
    the reference formulas applied to exactly the tallies the call returns."""
    P = S.multi_grain_like(n_photons_eq_th=500, tau_mid=20.0, pola=False)
    G = api.PhotonLoop(P)
    t = G.mc_photon_loop(1, 1, 500, 1.0e30, 1, False, **MIXED)
    T = G.temp_finale()
    T1 = G.temp_finale_nlte()
    G.close()
    assert np.allclose(T, S.temp_finale(P, t.xKJ_abs), rtol=5e-6)
    assert T.max() > 10 * P.T_min
    ref = _grain_temperatures(P, t, "nLTE")
    ks = np.arange(P.grain_RE_nLTE_start, P.grain_RE_nLTE_end + 1)
    C = P.C_abs_norm[ks - 1].astype(np.float64)
    E = (C @ (t.xJ_abs + P.J0).T) * P.L_packet_th / P.volume[None, :]
    ref = np.where(np.log(E) < P.log_E_em_1grain[:, :1], P.T_min, ref)
    assert T1.shape == ref.shape and np.allclose(T1, ref, rtol=5e-6)
    # an LTE-only handle has no per-grain state: loud error
    P2 = S.ref41_like(n_photons_eq_th=50, dark_zone=False, n_rad=20, nz=10, n_rad_in=3, tau_mid=10.0)
    G2 = api.PhotonLoop(P2)
    G2.mc_photon_loop(1, 1, 50)
    assert np.allclose(G2.temp_finale(), S.temp_finale(P2, G2.download().xKJ_abs), rtol=5e-6)
    with pytest.raises(api.McfostB200Error):
        G2.lib.mcfost_b200_temp_finale_nlte.argtypes = None
        G2._check(G2.lib.mcfost_b200_temp_finale_nlte(G2.h, np.zeros(4, np.float32).ctypes.data_as(__import__("ctypes").c_void_p)))
    G2.close()
