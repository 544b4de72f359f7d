"""CPU tests of the per-grain branches of the oracle: scattering method 1
(dust_transfer.f90:1291-1317) and nLTE / qRE re-emission (:1353-1395,
thermal_emission.f90:775-866, 1441-1514, 1953-2040).  Run with -m "not gpu"."""
import numpy as np
import pytest

from mcfost_b200 import synthetic as S
from oracle.binding import Oracle

MIXED = dict(lonly_LTE=0, lRE_nLTE=1, lnRE=1, lxJ_abs_step1=1)


def test_mixed_regimes_run_and_conserve_packets():
    P = S.multi_grain_like(n_photons_eq_th=40)
    O = Oracle(P)
    t = O.run(n_threads=1, xJ=True, n_photons2=40, **MIXED)
    st = t.stats
    assert st[0] == 128 * 40
    assert st[5] + st[6] == st[0]                 # every packet is killed or escapes
    assert st[4] > 0 and st[3] > 0                # absorptions and scatterings happened
    assert 0.0 < t.E_abs_nRE[0] < st[4]           # some energy went to grains out of equilibrium
    # absorptions happened on grains of both per-grain regimes and raised their temperature index
    assert (t.xT_ech_1grain > 2).any() and (t.xT_ech_1grain_nRE > 2).any()
    assert t.xT_ech_1grain.shape == (3, P.n_cells) and t.xT_ech_1grain.max() <= P.n_T
    # reproducible
    t2 = O.run(n_threads=1, xJ=True, n_photons2=40, **MIXED)
    assert np.array_equal(t.xKJ_abs, t2.xKJ_abs) and np.array_equal(t.xT_ech_1grain_nRE, t2.xT_ech_1grain_nRE)


def test_mixed_branch_with_only_lte_probability_is_the_lte_run():
    """Proba_abs_RE_LTE = 1 and proba_abs_RE = 1 send every absorption through im_reemission_LTE with the
    same rand / rand2 words: the tallies must equal the lonly_LTE run bit for bit."""
    P = S.multi_grain_like(n_photons_eq_th=30)
    P.Proba_abs_RE_LTE[:] = 1.0
    P.Proba_abs_RE_LTE_p_nLTE[:] = 1.0
    P.proba_abs_RE[:] = 1.0
    O = Oracle(P)
    a = O.run(n_threads=1, xJ=True, n_photons2=30, **MIXED)
    b = O.run(n_threads=1, xJ=True, n_photons2=30, lonly_LTE=1, lxJ_abs_step1=1)
    assert np.array_equal(a.xKJ_abs, b.xKJ_abs) and np.array_equal(a.xJ_abs, b.xJ_abs)
    assert np.array_equal(a.sed, b.sed) and a.E_abs_nRE[0] == 0.0


@pytest.mark.parametrize("low_mem", [0, 1])
def test_nlte_grain_selection_branches(low_mem):
    """select_absorbing_grain (low memory) reaches every nLTE grain size.  The kabs_nLTE_CDF bisection
    starts with kmin = grain_RE_nLTE_start and returns kmax (thermal_emission.f90:798-811), so it can
    never return the FIRST nLTE grain: a reference quirk the oracle reproduces."""
    P = S.multi_grain_like(n_photons_eq_th=60, tau_mid=5.0)
    O = Oracle(P)
    kw = dict(lonly_LTE=0, lonly_nLTE=1, lRE_nLTE=1, lxJ_abs_step1=1, n_photons2=60)
    t = O.run(n_threads=1, xJ=True, low_mem_th_emission_nLTE=low_mem, **kw)
    assert t.stats[5] + t.stats[6] == t.stats[0]
    touched = (t.xT_ech_1grain > 2).sum(axis=1)
    if low_mem:
        assert (touched > 0).all()
    else:
        assert touched[0] == 0 and (touched[1:] > 0).all()


def test_nlte_grain_temperature_is_analytic_in_the_thin_limit():
    """Optically thin disk, lonly_nLTE: a grain of size k at distance d from the star balances
    sum_l C_abs(k,l) 4 pi B_l(T_k) = sum_l C_abs(k,l) pi B_l(T_star) (R_star/d)^2.
    (a) the energy im_reemission_NLTE derives from the final xJ_abs gives that temperature through
    log_E_em_1grain; (b) xT_ech_1grain, set from the RUNNING xJ_abs at the grain's last absorption, is a
    tabulated temperature index at or below the one of the final energy (+1 grid step)."""
    P = S.multi_grain_like(n_photons_eq_th=1500, tau_mid=0.02, n_rad=10, nz=6, n_rad_in=2, n_T=120)
    O = Oracle(P)
    t = O.run(n_threads=0, xJ=True, lonly_LTE=0, lonly_nLTE=1, lRE_nLTE=1, lxJ_abs_step1=1, n_photons2=1500)
    wl = P.tab_lambda * 1e-6; dwl = P.tab_delta_lambda * 1e-6

    def planck(T):
        x = S.THERMAL_CONST / (T * wl)
        return np.where(x < 500, 1.0 / (wl ** 5 * np.expm1(np.minimum(x, 500))), 0.0) * dwl
    Rs = P.star_xyzr[3, 0]
    ks = np.arange(P.grain_RE_nLTE_start, P.grain_RE_nLTE_end + 1)
    Tgrid = P.tab_Temp.astype(np.float64)
    n_checked = 0
    err = []
    for ic in range(P.n_cells):
        d2 = P.r_grid[ic] ** 2 + P.z_grid[ic] ** 2
        for j, k in enumerate(ks):
            C = P.C_abs_norm[k - 1].astype(np.float64)
            absorbed = np.sum(C * np.pi * planck(P.star_T[0])) * Rs ** 2 / d2
            emitted = np.array([np.sum(C * 4 * np.pi * planck(T)) for T in Tgrid])
            T_an = np.exp(np.interp(np.log(absorbed), np.log(emitted), np.log(Tgrid)))
            log_E_abs = np.log(np.sum(C * (t.xJ_abs[ic] + P.J0[ic])) * P.L_packet_th / P.volume[ic])
            T_mc = np.exp(np.interp(log_E_abs, P.log_E_em_1grain[j], np.log(Tgrid)))
            err.append(T_mc / T_an - 1.0)
            Ti = t.xT_ech_1grain[j, ic]
            if Ti > 2:
                i_final = int(np.searchsorted(P.log_E_em_1grain[j], log_E_abs)) + 1      # first T index above the final energy
                assert Ti <= i_final + 1, (ic, k, Ti, i_final)
                n_checked += 1
    err = np.array(err)
    assert np.median(np.abs(err)) < 0.03 and np.percentile(np.abs(err), 90) < 0.10
    assert n_checked > 20


def test_method1_with_identical_grains_matches_method2_statistically():
    """If every grain has the population's phase function, choosing the grain first (method 1) samples the
    same scattering law as method 2: the temperature structure agrees within Monte Carlo noise."""
    P = S.multi_grain_like(n_photons_eq_th=300, tau_mid=20.0, n_rad=12, nz=8, n_rad_in=3, pola=False)
    P.tab_g[:] = P.tab_g_pos.reshape(1, -1)
    O = Oracle(P)
    kw = dict(n_photons2=300, lmethod_aniso1=0, lonly_LTE=1)
    a = O.run(n_threads=0, lscattering_method1=1, **kw)
    b = O.run(n_threads=0, lscattering_method1=0, **kw)
    assert a.stats[5] + a.stats[6] == a.stats[0]
    big = b.xKJ_abs > np.percentile(b.xKJ_abs, 50)
    rel = np.abs(a.xKJ_abs[big] / b.xKJ_abs[big] - 1.0)
    assert np.median(rel) < 0.1
    assert abs(a.stats[3] / b.stats[3] - 1.0) < 0.05          # same number of scatterings within noise


@pytest.mark.parametrize("low_mem", [0, 1])
def test_method1_grain_choice_follows_ksca_cdf(low_mem):
    """Method 1, Mie tables, Stokes on, both grain-selection branches: runs, conserves packets and gives
    the same scattering count as the other branch within noise."""
    P = S.multi_grain_like(n_photons_eq_th=100, tau_mid=10.0)
    O = Oracle(P)
    kw = dict(n_photons2=100, lscattering_method1=1, lmethod_aniso1=1, lsepar_pola=1)
    t = O.run(n_threads=1, low_mem_scattering=low_mem, **kw)
    assert t.stats[5] + t.stats[6] == t.stats[0]
    pol = np.abs(t.sed_q).sum() + np.abs(t.sed_u).sum()
    assert pol > 0.0                                          # the per-grain Mueller matrix polarises the packets
    o = O.run(n_threads=1, low_mem_scattering=1 - low_mem, **kw)
    assert abs(t.stats[3] / o.stats[3] - 1.0) < 0.1


def test_variable_dust_mixed_regimes():
    P = S.multi_grain_like(n_photons_eq_th=30, variable=True, pola=False)
    O = Oracle(P)
    t = O.run(n_threads=1, xJ=True, n_photons2=30, low_mem_th_emission_nLTE=1, **MIXED)
    assert t.stats[5] + t.stats[6] == t.stats[0]
    assert (t.xT_ech_1grain > 2).any()
    t1 = O.run(n_threads=1, n_photons2=30, lscattering_method1=1, lmethod_aniso1=0, low_mem_scattering=1)
    assert t1.stats[5] + t1.stats[6] == t1.stats[0]


def test_low_memory_lte_emission_matches_the_cell_cdf_statistically():
    """low_mem_th_emission draws the absorbing LTE grain and bisects its own CDF; summed over grains with the
    absorption weights this is the cell's kdB_dT_CDF, so the temperature structure and the emergent spectrum agree
    with the high-memory branch within noise."""
    P = S.multi_grain_like(n_photons_eq_th=300, tau_mid=20.0, n_rad=12, nz=8, n_rad_in=3, pola=False)
    O = Oracle(P)
    a = O.run(n_threads=0, n_photons2=300, lonly_LTE=1, low_mem_th_emission=1)
    b = O.run(n_threads=0, n_photons2=300, lonly_LTE=1, low_mem_th_emission=0)
    assert a.stats[5] + a.stats[6] == a.stats[0]
    assert abs(a.stats[4] / b.stats[4] - 1.0) < 0.05
    big = b.xKJ_abs > np.percentile(b.xKJ_abs, 50)
    assert np.median(np.abs(a.xKJ_abs[big] / b.xKJ_abs[big] - 1.0)) < 0.1
    na, nb = a.n_phot_sed.sum(axis=(1, 2)), b.n_phot_sed.sum(axis=(1, 2))
    m = (na + nb) > 200
    z = (na[m] - nb[m]) / np.sqrt(na[m] + nb[m])
    assert np.mean(np.abs(z) < 4) > 0.9


def test_hot_spot_and_weighted_emission_scale_the_packet_energy():
    P = S.multi_grain_like(n_photons_eq_th=50, tau_mid=5.0, pola=False)
    S.repartition_energie(P, Tdust=np.full(P.n_cells, 150.0))        # warm disk: some packets are emitted by the dust
    lam = int(np.argmax(P.tab_lambda > 30.0)) + 1
    assert 0.0 < P.frac_E_stars[lam - 1] < 1.0
    O = Oracle(P)
    kw = dict(letape_th=0, lmono=1, lambda_in=lam, p_lambda_in=lam, n_photons2=10 ** 9, n_phot_lim=200.0)
    base = O.run(n_threads=1, **kw)
    # hot spot covering half the star, 2x hotter: the star-light energy goes up, packet counts do not change
    spot = O.run(n_threads=1, lspot=1, T_spot=12000.0, surf_fraction_spot=0.75, theta_spot=0.0, phi_spot=0.0,
                 star1_T=float(P.star_T[0]), tab_lambda=P.tab_lambda, **kw)
    assert np.array_equal(spot.n_phot_sed, base.n_phot_sed)
    assert spot.sed_star.sum() > 1.2 * base.sed_star.sum() and np.isclose(spot.sed_disk.sum(), base.sed_disk.sum())
    # weighted emission: every disk packet carries correct_E_emission of its cell
    P.correct_E_emission = np.full(P.n_cells, 0.25)
    O.set_emission(P)
    w = O.run(n_threads=1, lweight_emission=1, **kw)
    assert np.array_equal(w.n_phot_sed, base.n_phot_sed)
    assert np.isclose(w.sed_disk.sum(), 0.25 * base.sed_disk.sum(), rtol=1e-12) and base.sed_disk.sum() > 0
    assert np.isclose(w.sed_star.sum(), base.sed_star.sum(), rtol=1e-12)


def test_temp_finale_and_temp_finale_nlte_of_the_oracle():
    """Temp_finale / Temp_finale_nLTE restated in the oracle against their numpy restatements on the same tallies."""
    P = S.multi_grain_like(n_photons_eq_th=200, tau_mid=20.0, pola=False)
    O = Oracle(P)
    t = O.run(n_threads=0, xJ=True, n_photons2=200, **MIXED)
    T = O.temp_finale()
    ref = S.temp_finale(P, t.xKJ_abs)
    assert T.shape == (P.n_cells,) and np.allclose(T, ref, rtol=2e-6)
    T1 = O.temp_finale_nlte()
    ks = np.arange(P.grain_RE_nLTE_start, P.grain_RE_nLTE_end + 1)
    C = P.C_abs_norm[ks - 1].astype(np.float64)
    E = (C @ (t.xJ_abs + P.J0).T) * P.L_packet_th / P.volume[None, :]
    lt = np.log(P.tab_Temp.astype(np.float64))
    for j in range(len(ks)):
        ref1 = np.exp(np.interp(np.log(E[j]), P.log_E_em_1grain[j], lt))
        ref1 = np.where(np.log(E[j]) < P.log_E_em_1grain[j, 0], P.T_min, ref1)
        assert np.allclose(T1[j], ref1, rtol=2e-6), j
