"""G2 (Pascucci_3.0.para, the 2D disk benchmark of Pascucci et al. 2004) on the oracle: the optically thin case has a
closed-form answer, which anchors the oracle's thermal step to something outside this repository."""
import numpy as np

from mcfost_b200 import synthetic as S
from oracle.binding import Oracle


def thin_limit_tally(P, n_packets):
    """Expected xKJ_abs per cell when nothing attenuates the star: N_lambda packets of wavelength lambda leave an
    isotropic point source, the summed path length inside a cell of volume V at distance d is N V / (4 pi d^2)."""
    spec = np.diff(np.asarray(P.spectre_emission_cumul))                 # packets per wavelength
    kabs = np.asarray(P.kappa_abs_LTE).reshape(-1)
    d2 = np.asarray(P.r_grid) ** 2 + np.asarray(P.z_grid) ** 2
    return n_packets * P.E_paquet * float(np.sum(spec * kabs)) * np.asarray(P.volume) / (4.0 * np.pi * d2)


def test_optically_thin_pascucci_disk_matches_radiative_equilibrium():
    n2 = 4000
    P = S.pascucci_like(tau_V=0.1, n_photons_eq_th=n2)
    assert P.n_lambda == 61 and P.n_cells == 7000 and abs(P.tab_lambda[0] - 0.110662) < 0.02
    O = Oracle(P, fast=True)
    t = O.run(n_threads=0, n_photons2=n2, lisotropic=1)
    assert t.stats[5] + t.stats[6] == t.stats[0] == 128 * n2
    T_mc = O.temp_finale()
    T_thin = S.temp_finale(P, thin_limit_tally(P, 128 * n2))
    # cells well inside the grid, away from the inner rim (finite star) and with enough packets
    sel = (np.asarray(P.r_grid) > 2.0) & (np.asarray(P.r_grid) < 800.0) & (t.xKJ_abs > 0) & (T_thin > 2.0)
    rel = (T_mc[sel] - T_thin[sel]) / T_thin[sel]
    # tau_V = 0.1 along the midplane: the stellar flux is attenuated by <= 10 % (2.5 % in T), scattered light adds a few %
    assert sel.sum() > 4000
    assert abs(np.median(rel)) < 0.02, np.median(rel)
    assert np.percentile(np.abs(rel), 90) < 0.06
    # the midplane profile follows the thin-limit slope
    mid = sel & (np.asarray(P.cell_map_j[:P.n_cells]) == 1)
    slope_mc = np.polyfit(np.log(np.asarray(P.r_grid)[mid]), np.log(T_mc[mid]), 1)[0]
    slope_thin = np.polyfit(np.log(np.asarray(P.r_grid)[mid]), np.log(T_thin[mid]), 1)[0]
    assert abs(slope_mc - slope_thin) < 0.02, (slope_mc, slope_thin)


def test_optical_depth_series_is_monotonic():
    """tau_V = 0.1, 1, 10, 100: the midplane gets colder and more packets interact as the disk gets thicker"""
    out = []
    for tau in (0.1, 1.0, 10.0, 100.0):
        P = S.pascucci_like(tau_V=tau, n_photons_eq_th=1000)
        O = Oracle(P, fast=True)
        t = O.run(n_threads=0, n_photons2=1000, lisotropic=1)
        T = O.temp_finale()
        mid = (np.asarray(P.cell_map_j[:P.n_cells]) == 1) & (np.asarray(P.r_grid) > 100.0) & (np.asarray(P.r_grid) < 300.0)
        out.append((t.stats[2] / t.stats[0], float(np.mean(T[mid]))))
        # tau_V is what the generator promises: the radial midplane optical depth at 0.55 um
        ic = np.array([1], np.int32)
        r = O.optical_length_tot(P.lambda_seuil, np.array([P.r_lim[0] * 1.0000001]), np.zeros(1), np.array([1e-9]), np.ones(1), np.zeros(1), np.zeros(1), ic)
        assert abs(r["tau_tot"][0] / tau - 1) < 0.02
    inter, Tmid = zip(*out)
    assert all(a < b for a, b in zip(inter, inter[1:]))
    assert all(a > b for a, b in zip(Tmid[1:], Tmid[2:]))
