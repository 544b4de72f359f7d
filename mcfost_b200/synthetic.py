"""Synthetic inputs for the photon-packet path.

The reference's own inputs (MCFOST_UTILS optical constants, stellar spectra)
are not available offline, so the benchmark / parity problems are generated
here: the *grid* and the *thermal / emission tables* restate the
reference's setup routines so that the arrays handed across the C ABI have
exactly the layout and meaning the Fortran side would hand over; the *dust
optics* are a documented analytic stand-in.

This module is setup code -- it is on neither side of the parity check (both
the CUDA path and the oracle consume the arrays it produces).

Restated reference routines (paths relative to reference src/):
  * define_cylindrical_grid            cylindrical_grid.f90:183-676
  * disk density (power-law Gaussian)  density.f90:58-200
  * init_lambda                        wavelengths.f90:22-71
  * init_tab_Temp                      Temperature.f90:23-39
  * init_reemission                    thermal_emission.f90:404-644
  * repartition_energie                thermal_emission.f90:1771-1949
  * repartition_wl_em                  thermal_emission.f90:315-360
  * calc_local_scattering_matrices     dust_prop.f90:1037-1243 (normalisations)
  * define_dark_zone                   optical_depth.f90:1425-1651
"""
from __future__ import annotations

from dataclasses import dataclass, field
from types import SimpleNamespace

import numpy as np

from .abi import MCB_GRID_CYL, MCB_GRID_SPH, MCB_GRID_VORONOI, NANG_SCATT

# constants.f90
PI = 3.141592653589793238462643383279502884197
HP = 6.626070040e-34
KB = 1.38064852e-23
C_LIGHT = 299792458.0
THERMAL_CONST = float(np.float32(C_LIGHT * HP / KB))   # `real, parameter :: thermal_const`
RSUN_TO_AU = 6.957e8 / 149597870700.0
CUTOFF = 7.0                                           # parameters.f90:111


@dataclass
class DiskZone:
    """disk_zone_type fields used by the grid / density builders."""
    rin: float = 1.0
    rout: float = 300.0
    edge: float = 0.0
    sclht: float = 10.0      # scale height at rref [AU]
    rref: float = 100.0
    exp_beta: float = 1.125  # flaring exponent
    surf: float = -0.5       # surface-density exponent
    dust_mass: float = 1.0e-3

    @property
    def rmin(self):
        return self.rin - 5.0 * self.edge


class Problem(SimpleNamespace):
    """Bag of arrays named exactly like the reference's module variables."""


# ---------------------------------------------------------------------------
# grid
# ---------------------------------------------------------------------------
def _radial_grid(n_rad, n_rad_in, zones):
    """tab_r(1:n_rad+1), cylindrical_grid.f90:258-368 (single region, log grid)."""
    rmin = min(z.rmin for z in zones)
    rmax = max(z.rout for z in zones)
    n_rad_in = max(n_rad_in, 1)
    tab_r = np.zeros(n_rad + 2)  # 1-based
    R0 = rmin
    tab_r[1] = R0
    ln_delta_r = (1.0 / float(n_rad - n_rad_in + 1)) * np.log(rmax / R0)
    delta_r = np.exp(ln_delta_r)
    puiss = 0.0
    for z in zones:
        p = 1 + z.surf - z.exp_beta
        if p > puiss:
            puiss = p
    if puiss == 0.0:
        for i in range(2, 2 + n_rad_in):
            tab_r[i] = np.exp(np.log(R0) - (np.log(R0) - np.log(R0 * delta_r))
                              * (2.0 ** (i - 1) - 1.0) / (2.0 ** n_rad_in - 1.0))
    else:
        for i in range(2, 2 + n_rad_in):
            tab_r[i] = (R0 ** puiss - (R0 ** puiss - (R0 * delta_r) ** puiss)
                        * (2.0 ** (i - 1 + 1) - 1.0) / (2.0 ** (n_rad_in + 1) - 1.0)) ** (1.0 / puiss)
    for i in range(2 + n_rad_in, n_rad + 2):
        tab_r[i] = tab_r[i - 1] * delta_r
    return rmin, rmax, tab_r


def cell_numbering(n_rad, nz, n_az, l3D):
    """build_cylindrical_cell_mapping (cylindrical_grid.f90:45-179) in numpy:
    returns cell_map_i/j/k (1-based ids, length ntot2) -- handed over the ABI so
    the library can verify its analytic numbering."""
    j_start = -nz if l3D else 1
    ci, cj, ck = [], [], []
    for k in range(1, n_az + 1):
        for j in range(j_start, nz + 1):
            if j == 0:
                continue
            for i in range(1, n_rad + 1):
                ci.append(i); cj.append(j); ck.append(k)
    jstart2 = min(1, j_start) - 1
    jend2 = nz + 1
    for k in range(1, n_az + 1):
        for j in (jstart2, jend2):
            for i in range(0, n_rad + 2):
                ci.append(i); cj.append(j); ck.append(k)
    for k in range(1, n_az + 1):
        for j in range(j_start, nz + 1):
            if j == 0:
                continue
            for i in (0, n_rad + 1):
                ci.append(i); cj.append(j); ck.append(k)
    return (np.array(ci, np.int32), np.array(cj, np.int32), np.array(ck, np.int32))


def cylindrical_grid(n_rad=100, nz=70, n_az=1, n_rad_in=20, zones=None, l3D=False):
    zones = zones or [DiskZone()]
    P = Problem()
    P.kind, P.l3D = MCB_GRID_CYL, int(l3D)
    P.n_rad, P.nz, P.n_az = n_rad, nz, n_az
    P.n_cells = n_rad * nz * n_az * (2 if l3D else 1)
    rmin, rmax, tab_r = _radial_grid(n_rad, n_rad_in, zones)
    P.Rmax2 = rmax * rmax
    tab_r2 = tab_r * tab_r
    tab_r3 = tab_r2 * tab_r
    r_lim = np.zeros(n_rad + 1); r_lim_2 = np.zeros(n_rad + 1); r_lim_3 = np.zeros(n_rad + 1)
    r_lim[0], r_lim_2[0], r_lim_3[0] = rmin, rmin ** 2, rmin ** 3
    for i in range(1, n_rad + 1):
        r_lim[i], r_lim_2[i], r_lim_3[i] = tab_r[i + 1], tab_r2[i + 1], tab_r3[i + 1]
    P.r_lim, P.r_lim_2, P.r_lim_3 = r_lim, r_lim_2, r_lim_3
    # zmax, cell_height, z_lim  (:416-494)
    zmax = np.zeros(n_rad)
    rcyl_c = np.zeros(n_rad)
    for i in range(1, n_rad + 1):
        rcyl = 0.5 * (r_lim[i] + r_lim[i - 1])
        rcyl_c[i - 1] = rcyl
        H = 0.0
        for z in zones:
            if z.rmin < rcyl < z.rout:
                hz = z.sclht * (rcyl / z.rref) ** z.exp_beta
                H = max(H, hz)
        zmax[i - 1] = CUTOFF * H
    for i in range(n_rad):          # interpolation between zones (:433-455)
        if zmax[i] < np.finfo(np.float32).tiny:
            lo = max(ii for ii in range(i) if zmax[ii] > 0)
            hi = min(ii for ii in range(i + 1, n_rad) if zmax[ii] > 0)
            frac = (np.log(rcyl_c[i]) - np.log(rcyl_c[lo])) / (np.log(rcyl_c[hi]) - np.log(rcyl_c[lo]))
            zmax[i] = np.exp(np.log(zmax[hi]) * frac + np.log(zmax[lo]) * (1.0 - frac))
    cell_height = zmax / float(np.float32(nz))
    z_lim = np.zeros((n_rad, nz + 2), order="F")
    for j in range(1, nz + 1):
        z_lim[:, j - 1] = (float(j) - 1.0) * cell_height
    z_lim[:, nz] = zmax
    z_lim[:, nz + 1] = float(np.float32(1.0e30))
    P.zmax, P.z_lim, P.zmaxmax = zmax, z_lim, float(zmax.max())
    # volumes (:479-491, 625): note the routine's local fp32 pi (:191)
    pi32 = float(np.float32(3.1415926535))
    V = np.zeros((n_rad, nz))
    for i in range(1, n_rad + 1):
        if (tab_r2[i + 1] - tab_r2[i]) > 1.0e-6 * tab_r2[i]:
            dr2 = 2.0 * pi32 * (tab_r2[i + 1] - tab_r2[i])
        else:
            dr2 = 4.0 * pi32 * rcyl_c[i - 1] * (tab_r[i + 1] - tab_r[i])
        V[i - 1, :] = dr2 * cell_height[i - 1]
    z_c = z_lim[:, :nz] + 0.5 * cell_height[:, None]
    if l3D:
        V = V * 0.5 / float(np.float32(n_az))
        # tan_phi_lim (:586-600) in fp32 like the reference
        d_phi = np.float32(2.0) * np.float32(3.1415926535) / np.float32(n_az)
        tan_phi = np.zeros(n_az); cos_phi = np.zeros(n_az); sin_phi = np.zeros(n_az)
        for k in range(1, n_az + 1):
            phi = np.float32(d_phi * np.float32(k))
            m = np.float32(np.mod(np.float32(phi - np.float32(0.5) * np.float32(3.1415926535)), np.float32(3.1415926535)))
            tan_phi[k - 1] = 1.0e300 if abs(m) < 1.0e-6 else float(np.tan(np.float32(phi)))
            # cos_phi_lim / sin_phi_lim (:592-598), read by distance_to_closest_wall_* only
            cos_phi[k - 1] = 0.0 if abs(m) < 1.0e-6 else float(np.cos(np.float32(phi)))
            sin_phi[k - 1] = 1.0e300 if abs(m) < 1.0e-6 else float(np.sin(np.float32(phi)))
        P.tan_phi_lim = tan_phi
        P.cos_phi_lim, P.sin_phi_lim = cos_phi, sin_phi
    else:
        P.tan_phi_lim = np.zeros(max(n_az, 1))
        P.cos_phi_lim = np.zeros(max(n_az, 1)); P.sin_phi_lim = np.zeros(max(n_az, 1))
    ci, cj, ck = cell_numbering(n_rad, nz, n_az, l3D)
    P.cell_map_i, P.cell_map_j, P.cell_map_k = ci, cj, ck
    P.n_cells_tot = len(ci)
    nc = P.n_cells
    ii, jj, kk = ci[:nc] - 1, cj[:nc], ck[:nc]
    ja = np.abs(jj) - 1
    P.volume = V[ii, ja].copy()
    P.r_grid = rcyl_c[ii].copy()
    P.z_grid = np.where(jj > 0, z_c[ii, ja], -z_c[ii, ja])
    dphi = 2.0 * PI / n_az
    P.phi_grid = (dphi * (kk - 0.5)) if l3D else np.zeros(nc)
    P.tan_theta_lim = None; P.theta_lim = None
    P.zones = zones
    return P


def spherical_grid(n_rad=60, nz=30, n_az=1, n_rad_in=5, rin=10.0, rout=200.0, l3D=False):
    """define_cylindrical_grid, lspherical branch (:496-580), uniform in cos."""
    zone = DiskZone(rin=rin, rout=rout, surf=-1.0, exp_beta=1.0)
    P = Problem()
    P.kind, P.l3D = MCB_GRID_SPH, int(l3D)
    P.n_rad, P.nz, P.n_az = n_rad, nz, n_az
    P.n_cells = n_rad * nz * n_az * (2 if l3D else 1)
    rmin, rmax, tab_r = _radial_grid(n_rad, n_rad_in, [zone])
    P.Rmax2 = rmax * rmax
    r_lim = np.zeros(n_rad + 1)
    r_lim[0] = rmin
    r_lim[1:] = tab_r[2:n_rad + 2]
    P.r_lim, P.r_lim_2, P.r_lim_3 = r_lim, r_lim ** 2, r_lim ** 3
    P.r_lim_2[1:] = (tab_r * tab_r)[2:n_rad + 2]
    P.r_lim_3[1:] = (tab_r * tab_r * tab_r)[2:n_rad + 2]
    w_lim = np.zeros(nz + 1); theta_lim = np.zeros(nz + 1); tan_theta_lim = np.zeros(nz + 1)
    tan_theta_lim[0] = 1.0e-10
    w_lim[nz] = 1.0; theta_lim[nz] = PI / 2.0; tan_theta_lim[nz] = 1.0e30
    for j in range(1, nz):
        w = float(j) / float(nz)
        w_lim[j] = w
        c = np.sqrt(1.0 - w * w)
        tan_theta_lim[j] = w / c
        theta_lim[j] = np.arctan(tan_theta_lim[j])
    P.tan_theta_lim, P.theta_lim, P.w_lim = tan_theta_lim, theta_lim, w_lim
    dcos = 1.0 / float(np.float32(nz))
    pi32 = float(np.float32(3.1415926535))
    V = np.zeros((n_rad, nz)); rg = np.zeros((n_rad, nz)); zg = np.zeros((n_rad, nz))
    tab_r3 = tab_r ** 3
    for i in range(1, n_rad + 1):
        rsph = np.sqrt(r_lim[i] * r_lim[i - 1])
        for j in range(1, nz + 1):
            w = 0.5 * (w_lim[j] + w_lim[j - 1])
            rg[i - 1, j - 1] = rsph * np.sqrt(1.0 - w * w)
            zg[i - 1, j - 1] = rsph * w
        Vi = 4.0 / 3.0 * pi32 * (tab_r3[i + 1] - tab_r3[i])
        V[i - 1, :] = Vi * dcos
    if l3D:
        V = V * 0.5 / float(np.float32(n_az))
        d_phi = np.float32(2.0) * np.float32(3.1415926535) / np.float32(n_az)
        tan_phi = np.zeros(n_az); cos_phi = np.zeros(n_az); sin_phi = np.zeros(n_az)
        for k in range(1, n_az + 1):
            phi = np.float32(d_phi * np.float32(k))
            m = np.float32(np.mod(np.float32(phi - np.float32(0.5) * np.float32(3.1415926535)), np.float32(3.1415926535)))
            tan_phi[k - 1] = 1.0e300 if abs(m) < 1.0e-6 else float(np.tan(np.float32(phi)))
            # cos_phi_lim / sin_phi_lim (:592-598), read by distance_to_closest_wall_* only
            cos_phi[k - 1] = 0.0 if abs(m) < 1.0e-6 else float(np.cos(np.float32(phi)))
            sin_phi[k - 1] = 1.0e300 if abs(m) < 1.0e-6 else float(np.sin(np.float32(phi)))
        P.tan_phi_lim = tan_phi
        P.cos_phi_lim, P.sin_phi_lim = cos_phi, sin_phi
    else:
        P.tan_phi_lim = np.zeros(max(n_az, 1))
        P.cos_phi_lim = np.zeros(max(n_az, 1)); P.sin_phi_lim = np.zeros(max(n_az, 1))
    ci, cj, ck = cell_numbering(n_rad, nz, n_az, l3D)
    P.cell_map_i, P.cell_map_j, P.cell_map_k = ci, cj, ck
    P.n_cells_tot = len(ci)
    nc = P.n_cells
    ii, jj, kk = ci[:nc] - 1, cj[:nc], ck[:nc]
    ja = np.abs(jj) - 1
    P.volume = V[ii, ja].copy()
    P.r_grid = rg[ii, ja].copy()
    P.z_grid = np.where(jj > 0, zg[ii, ja], -zg[ii, ja])
    P.phi_grid = (2.0 * PI / n_az * (kk - 0.5)) if l3D else np.zeros(nc)
    P.z_lim = np.zeros((n_rad, nz + 2), order="F"); P.zmax = np.ones(n_rad); P.zmaxmax = 0.0
    P.zones = [zone]
    return P


# ---------------------------------------------------------------------------
# physics tables
# ---------------------------------------------------------------------------
def init_lambda(n_lambda=50, lambda_min=0.1, lambda_max=3000.0):
    """wavelengths.f90:41-57 (tab_lambda is `real` in the reference)."""
    delta = np.exp((1.0 / n_lambda) * np.log(lambda_max / lambda_min))
    lam = np.zeros(n_lambda); lsup = np.zeros(n_lambda); linf = np.zeros(n_lambda)
    linf[0] = lambda_min; lam[0] = lambda_min * np.sqrt(delta); lsup[0] = lambda_min * delta
    for i in range(1, n_lambda):
        lam[i] = lam[i - 1] * delta
        lsup[i] = lsup[i - 1] * delta
        linf[i] = lsup[i - 1]
    return lam, lsup - linf


def init_tab_Temp(n_T=100, T_min=1.0, T_max=3000.0):
    """Temperature.f90:23-39."""
    delta_T = np.exp((1.0 / n_T) * np.log(T_max / T_min))
    t = np.zeros(n_T, np.float32)
    t[0] = T_min * np.sqrt(delta_T)
    for k in range(1, n_T):
        t[k] = np.float32(delta_T * t[k - 1])
    return t


def disk_density(P, zones):
    """Power-law Gaussian disk, density.f90:147-185 evaluated at cell centres,
    normalised to the zone dust mass (arbitrary units: only ratios are used)."""
    rho = np.zeros(P.n_cells)
    for z in zones:
        r, zz = P.r_grid, P.z_grid
        fact = (r / z.rref) ** (z.surf - z.exp_beta)
        coeff = 2.0 * (r / z.rref) ** (2 * z.exp_beta)
        d = fact * np.exp(-((zz / z.sclht) ** 2) / coeff)
        d = np.where((r > z.rout) | (r < z.rmin), 0.0, d)
        m = float(np.sum(d * P.volume))
        rho += d * (z.dust_mass / m)
    return rho


def synthetic_optics(lam, pola=True, isotropic=False):
    """Documented stand-in for Mie theory on Draine silicates:
    kappa_ext ~ 1/lambda beyond 1 um (flat below), albedo 0.5 -> 0 across
    1-100 um, HG asymmetry 0.6 -> 0, Mueller matrix = HG s11 x Rayleigh-like
    polarisation.  Returned per wavelength, unit extinction at lambda <= 1 um."""
    n = len(lam)
    kext = np.where(lam > 1.0, 1.0 / lam, 1.0)
    x = np.clip(np.log10(np.maximum(lam, 1.0)) / 2.0, 0.0, 1.0)
    albedo = 0.5 * (1.0 - x)
    g = 0.0 * lam if isotropic else 0.6 * (1.0 - x)
    theta = np.arange(NANG_SCATT + 1) * PI / NANG_SCATT
    mu = np.cos(theta)
    s11 = np.zeros((NANG_SCATT + 1, n))
    for l in range(n):
        s11[:, l] = (1.0 - g[l] ** 2) * (1.0 + g[l] ** 2 - 2.0 * g[l] * mu) ** (-1.5)
    pmax = 0.4
    s12_o = -pmax * (1.0 - mu * mu) / (1.0 + mu * mu)
    s22_o = np.ones_like(mu)
    s33_o = 2.0 * mu / (1.0 + mu * mu)
    s34_o = 0.1 * np.sin(theta) ** 2 * mu      # small circular-polarisation term so V is exercised
    s44_o = s33_o.copy()
    return kext, albedo, g, s11, (s12_o, s22_o, s33_o, s34_o, s44_o) if pola else None


def scattering_tables(P, s11, albedo, kappa, pola_tabs):
    """calc_local_scattering_matrices normalisations, dust_prop.f90:1141-1178,
    for p_n_cells = 1 and p_n_lambda_pos = n_lambda."""
    n_lambda = s11.shape[1]
    dtheta = PI / float(np.float32(NANG_SCATT))
    theta = np.arange(NANG_SCATT + 1, dtype=np.float64)
    theta = np.float32(theta).astype(np.float64) * dtheta
    prob = np.zeros((NANG_SCATT + 1, 1, n_lambda), np.float32, order="F")
    tab = np.zeros((NANG_SCATT + 1, 1, n_lambda), np.float32, order="F")
    for l in range(n_lambda):
        ksca = float(kappa[l] * albedo[l])
        s = s11[:, l].astype(np.float64)
        if ksca > float(np.finfo(np.float32).tiny):
            norm = np.sum(s[1:NANG_SCATT] * np.sin(theta[1:NANG_SCATT]) * dtheta)
            s = (s * ksca / norm).astype(np.float32)      # tab_s11_pos normalised to k_sca_tot
            p = np.zeros(NANG_SCATT + 1, np.float32)
            for a in range(2, NANG_SCATT + 1):
                p[a] = np.float32(p[a - 1] + np.float32(float(s[a]) * np.sin(theta[a]) * dtheta))
            p[1:] = np.float32(p[1:] + np.float32(ksca - float(p[NANG_SCATT])))
            p = np.float32(p / np.float32(ksca))
            prob[:, 0, l] = p
            tab[:, 0, l] = np.float32(s.astype(np.float64) * dtheta / (ksca * 2.0 * PI))
        else:
            prob[:, 0, l] = 1.0
            prob[0, 0, l] = 0.0
            tab[:, 0, l] = 1.0
    out = dict(prob_s11_pos=prob, tab_s11_pos=tab)
    names = ("tab_s12_o_s11_pos", "tab_s22_o_s11_pos", "tab_s33_o_s11_pos", "tab_s34_o_s11_pos", "tab_s44_o_s11_pos")
    for nm, t in zip(names, pola_tabs or [None] * 5):
        if t is None:
            out[nm] = None
        else:
            a = np.zeros((NANG_SCATT + 1, 1, n_lambda), np.float32, order="F")
            a[:, 0, :] = np.float32(t)[:, None]
            out[nm] = a
    return out


def init_reemission(P):
    """thermal_emission.f90:404-550 (high-memory LTE branch, no extra heating)."""
    n_T, n_lambda, pnc = P.n_T, P.n_lambda, P.p_n_cells
    cst_E = 2.0 * HP * C_LIGHT ** 2 * 4.0 * PI
    wl = P.tab_lambda * 1.0e-6
    dwl = P.tab_delta_lambda * 1.0e-6
    B = np.zeros((n_lambda, n_T)); dB = np.zeros((n_lambda, n_T))
    for t in range(n_T):
        cst = THERMAL_CONST / float(P.tab_Temp[t])
        cst_wl = cst / wl
        ok = cst_wl < 500.0
        ce = np.exp(np.where(ok, cst_wl, 1.0))
        b = np.where(ok, 1.0 / ((wl ** 5) * (ce - 1.0)) * dwl, 0.0)
        B[:, t] = b
        dB[:, t] = np.where(ok, b * cst_wl * ce / (ce - 1.0), 0.0)
    logQ = np.zeros((n_T, pnc), order="F")
    cdf = np.zeros((n_lambda, n_T, pnc), order="F")
    kabs = P.kappa_abs_LTE.reshape(pnc, n_lambda)
    tiny_dp = np.finfo(np.float64).tiny
    # vectorised over the cells; every cell keeps the reference's left-to-right summation order over lambda
    Qcool0 = np.zeros(pnc)
    for t in range(n_T):
        integ = np.zeros(pnc)
        for l in range(n_lambda):
            integ = integ + kabs[:, l] * B[l, t]
        Qcool = integ * cst_E
        if t == 0:
            Qcool0 = Qcool
        q = Qcool - Qcool0
        pos = q > tiny_dp
        logQ[t, :] = np.where(pos, np.log(np.where(pos, q, 1.0)), -1000.0)
        integ3 = np.cumsum(kabs * dB[None, :, t], axis=1)            # (pnc, n_lambda), sequential along lambda
        tot = integ3[:, n_lambda - 1]
        ok = tot > tiny_dp
        cdf[:, t, :] = np.where(ok[None, :], integ3.T / np.where(ok, tot, 1.0)[None, :], 0.0)
    P.log_Qcool_minus_extra_heating = logQ
    P.kdB_dT_CDF = cdf
    return P


def star_energy(P):
    """Blackbody branch of stars.f90:549-556 + :581-605 (E_stars, CDF_E_star)."""
    wl = P.tab_lambda * 1.0e-6
    n_stars = P.n_stars
    prob = np.zeros((P.n_lambda, n_stars))
    for i in range(n_stars):
        surface = 4.0 * PI * P.star_xyzr[3, i] ** 2
        cst_wl = THERMAL_CONST / (P.star_T[i] * wl)
        tiny32 = float(np.finfo(np.float32).tiny)
        prob[:, i] = np.where(cst_wl < 500.0, surface / ((wl ** 5) * (np.exp(np.minimum(cst_wl, 500.0)) - 1.0)), tiny32)
    cdf = np.zeros((P.n_lambda, n_stars + 1), np.float32, order="F")
    for i in range(n_stars):
        cdf[:, i + 1] = np.float32(cdf[:, i] + np.float32(prob[:, i]))
    P.E_stars = cdf[:, n_stars].astype(np.float64)        # `real` E_stars
    P.CDF_E_star = np.asfortranarray(cdf / cdf[:, n_stars:n_stars + 1])
    return P


def repartition_energie(P, Tdust=None):
    """thermal_emission.f90:1771-1949 (LTE only) for every wavelength, then
    repartition_wl_em (:315-360).  Tdust defaults to reset_temperature's 1 K."""
    nc, nl = P.n_cells, P.n_lambda
    Tdust = np.ones(nc) if Tdust is None else np.asarray(Tdust, np.float64)
    wl = P.tab_lambda * 1.0e-6
    cst_wl_max = float(np.float32(np.log(np.finfo(np.float32).max) - 1.0e-4))
    kabs = P.kappa_abs_LTE.reshape(P.p_n_cells, nl)
    dark = P.l_dark_zone.astype(bool)
    prob = np.zeros((nc + 1, nl), order="F")
    E_disk = np.zeros(nl)
    for l in range(nl):
        cst = THERMAL_CONST / (np.maximum(Tdust, 1e-300) * wl[l])
        k = kabs[:, l] if P.p_n_cells > 1 else kabs[0, l]
        ok = (~dark) & (Tdust >= float(np.finfo(np.float32).tiny)) & (cst < cst_wl_max)
        E = np.where(ok, 4.0 * k * P.kappa_factor * P.volume / ((wl[l] ** 5) * (np.exp(np.where(ok, cst, 1.0)) - 1.0)), 0.0)
        E_disk[l] = E.sum()
        c = np.concatenate(([0.0], np.cumsum(E)))
        prob[:, l] = c / c[nc] if c[nc] > np.finfo(np.float64).tiny else 0.0
    E_ISM = np.asarray(getattr(P, "E_ISM", np.zeros(nl)), np.float64)      # interstellar radiation field (stars.f90:607-640), optional
    tot = P.E_stars + E_disk + E_ISM
    P.E_disk = E_disk
    P.frac_E_stars = P.E_stars / tot
    P.frac_E_disk = (P.E_stars + E_disk) / tot
    P.prob_E_cell = prob
    dwl = P.tab_delta_lambda * 1.0e-6
    cum = np.concatenate(([0.0], np.cumsum(tot * dwl)))
    P.spectre_emission_cumul = cum / cum[nl]
    L_tot = 2.0 * PI * HP * C_LIGHT ** 2 * float(np.sum(tot * dwl))
    P.L_tot = L_tot
    P.L_packet_th = L_tot / float(np.float32(P.n_photons_loop) * np.float32(P.n_photons_eq_th))
    return P


def define_dark_zone(P, lambda_idx, tau_max=1500.0, physical_length=None):
    """optical_depth.f90:1425-1651 for the 2D cylindrical case.

    Step 4 of the reference shoots 11 rays per cell with physical_length();
    ``physical_length`` is a callable implementing that walk (the CUDA library's
    deterministic kernel in production, the oracle in CPU tests):
        physical_length(lambda_1based, x,y,z,u,v,w, icell, tau, dark) ->
            flag_sortie (bool array)
    With ``physical_length=None`` only steps 1-3 are applied and a cell is dark
    if it is deeper than tau_max radially (both ways) and vertically (a
    conservative superset check used for quick tests)."""
    n_rad, nz = P.n_rad, P.nz
    assert P.kind == MCB_GRID_CYL and not P.l3D
    kap = P.kappa.reshape(P.p_n_cells, P.n_lambda)[0, lambda_idx - 1] * P.kappa_factor   # (n_cells)
    cm = lambda i, j: (i - 1) + n_rad * (j - 1)      # 0-based id of real cell (i,j)
    ri_in, ri_out = n_rad, 1
    s = 0.0
    for i in range(1, n_rad + 1):
        s += kap[cm(i, 1)] * (P.r_lim[i] - P.r_lim[i - 1])
        if s > tau_max:
            ri_in = i
            break
    s = 0.0
    for i in range(n_rad, 0, -1):
        s += kap[cm(i, 1)] * (P.r_lim[i] - P.r_lim[i - 1])
        if s > tau_max:
            ri_out = i
            break
    if ri_out == n_rad:
        ri_out = n_rad - 1
    zj_sup = np.zeros(n_rad + 1, np.int64)
    for i in range(ri_in, ri_out + 1):
        s = 0.0
        for j in range(nz, 0, -1):
            s += kap[cm(i, j)] * (P.z_lim[i - 1, j] - P.z_lim[i - 1, j - 1])
            if s > tau_max:
                zj_sup[i] = j
                break
    dark = np.zeros(P.n_cells, np.int32)
    nb_angle = 11
    for i in range(max(ri_in, 2), ri_out + 1):
        top = int(zj_sup[i])
        if top < 1:
            continue
        if physical_length is None:
            for j in range(top, 0, -1):
                dark[cm(i, j)] = 1
            continue
        js = np.arange(top, 0, -1)
        ic = np.array([cm(i, j) for j in js])
        ang = np.float32(PI) * (np.arange(1, nb_angle + 1, dtype=np.float32) / np.float32(nb_angle + 1))   # `real :: angle`
        x0 = np.repeat(P.r_grid[ic], nb_angle); z0 = np.repeat(P.z_grid[ic], nb_angle)
        u0 = np.tile(np.cos(ang).astype(np.float64), len(js)); w0 = np.tile(np.sin(ang).astype(np.float64), len(js))
        y0 = np.zeros_like(x0); v0 = np.zeros_like(x0)
        icell = np.repeat(ic + 1, nb_angle).astype(np.int32)
        sortie = physical_length(lambda_idx, x0, y0, z0, u0, v0, w0, icell, np.full(len(x0), tau_max, np.float32), dark)
        sortie = np.asarray(sortie, bool).reshape(len(js), nb_angle)
        trapped = ~sortie.all(axis=1)
        if trapped.any():
            jtop = js[np.argmax(trapped)]            # first (highest) j with a non-exiting ray
            for jj in range(1, jtop + 1):
                dark[cm(i, jj)] = 1
    # region edges are never dark (:1640-1645)
    for j in range(1, nz + 1):
        dark[cm(1, j)] = 0
        dark[cm(n_rad, j)] = 0
    return dark


# ---------------------------------------------------------------------------
# the benchmark configurations
# ---------------------------------------------------------------------------
def cell_centres(P):
    """Cartesian centre of every cell, as compute_column builds it (optical_depth.f90:362-370)."""
    if P.kind == MCB_GRID_VORONOI:
        return P.vor_xyz[0].copy(), P.vor_xyz[1].copy(), P.vor_xyz[2].copy()
    return P.r_grid * np.cos(P.phi_grid), P.r_grid * np.sin(P.phi_grid), P.z_grid.copy()


def envelope_density(P, azimuthal=0.0):
    """Synthetic envelope for the spherical grids: rho ~ r^-1.5 (1 + cos^2(colatitude))
    with an optional m=1 azimuthal modulation so that phi walls matter in 3D."""
    rs = np.sqrt(P.r_grid ** 2 + P.z_grid ** 2)
    rho = rs ** -1.5 * (1.0 + (P.z_grid / rs) ** 2)
    if azimuthal:
        rho = rho * (1.0 + azimuthal * np.cos(P.phi_grid))
    return rho


def _finish(P, rho, n_lambda, n_T, tau_mid, pola, isotropic, n_photons_eq_th, star_T, star_R,
            lam_range=(0.1, 3000.0), lam_tau=0.81, optics=None):
    P.n_lambda, P.n_T = n_lambda, n_T
    P.tab_lambda, P.tab_delta_lambda = init_lambda(n_lambda, lam_range[0], lam_range[1])
    P.tab_lambda = np.float32(P.tab_lambda).astype(np.float64)          # `real` tables in the reference
    P.tab_delta_lambda = np.float32(P.tab_delta_lambda).astype(np.float64)
    P.T_min, P.T_max = 1.0, 3000.0
    P.tab_Temp = init_tab_Temp(n_T, P.T_min, P.T_max)
    P.n_photons_loop, P.n_photons_eq_th = 128, n_photons_eq_th
    # star at the origin (inside the inner edge => virtual cell (0,1,1))
    P.n_stars = 1
    P.star_xyzr = np.asfortranarray(np.array([[0.0], [0.0], [0.0], [star_R * RSUN_TO_AU]]))
    P.star_T = np.array([star_T])
    P.star_out_model = np.zeros(1, np.int32)
    # density and opacities (lvariable_dust = .false. => p_n_cells = 1)
    P.p_n_cells, P.p_n_lambda_pos = 1, n_lambda
    rho0 = rho[int(np.argmax(rho))]          # icell_not_empty stand-in: the densest cell
    P.kappa_factor = rho / rho0
    kext, albedo, g, s11, pol = (optics or synthetic_optics)(P.tab_lambda, pola=pola, isotropic=isotropic)
    # scale so that the radial midplane optical depth at lam_tau (0.81 um) is tau_mid
    l_seuil = int(np.argmax(P.tab_lambda > lam_tau)) + 1
    P.lambda_seuil = l_seuil
    if P.kind == MCB_GRID_CYL:
        mid = np.array([(i - 1) + P.n_rad * (0 if not P.l3D else P.nz) for i in range(1, P.n_rad + 1)])
    else:
        mid = np.arange(P.n_rad)
    col = float(np.sum(P.kappa_factor[mid] * (P.r_lim[1:] - P.r_lim[:-1])))
    k0 = tau_mid / (col * kext[l_seuil - 1])
    P.kappa = np.asfortranarray((k0 * kext).reshape(1, n_lambda))
    P.tab_albedo_pos = np.asfortranarray(np.float32(albedo).reshape(1, n_lambda))
    P.tab_g_pos = np.asfortranarray(np.float32(g).reshape(1, n_lambda))
    P.kappa_abs_LTE = np.asfortranarray(P.kappa * (1.0 - P.tab_albedo_pos.astype(np.float64)))
    for k, v in scattering_tables(P, s11, albedo, P.kappa[0], pol).items():
        setattr(P, k, v)
    init_reemission(P)
    star_energy(P)
    P.l_dark_zone = np.zeros(P.n_cells, np.int32)
    P.E_paquet = 1.0
    P.R_ISM = 0.0
    P.centre_ISM = (0.0, 0.0, 0.0)
    return P


def locate_stars(P, index_cell):
    """stars_cell_indices (stars.f90:789-808): star(:)%icell via index_cell.
    ``index_cell(x,y,z) -> 1-based ids`` comes from whichever side is being set up."""
    xyz = P.star_xyzr
    P.star_icell = np.asarray(index_cell(xyz[0].copy(), xyz[1].copy(), xyz[2].copy()), np.int32)
    return P


def star_icell_analytic(P):
    """For a star at the origin inside r_lim(0): the virtual cell (0,1,1) =
    first virtual id of the numbering (cylindrical_grid.f90:123-141)."""
    nj = 2 * P.nz if P.l3D else P.nz
    if P.l3D:
        # j = -nz-1 row comes first, then j = nz+1 row: (0, 1, 1) is in neither; it is in
        # the second virtual block (i in {0, n_rad+1}, j over real rows)
        base = P.n_cells + 2 * (P.n_rad + 2) * P.n_az
        return base + 2 * P.nz + 1          # j=-nz..-1 (2 ids each) then j=1, i=0
    base = P.n_cells + 2 * (P.n_rad + 2) * P.n_az
    return base + 1                         # j=1, i=0


def ref41_like(n_photons_eq_th=1000, tau_mid=1.0e5, pola=True, n_rad=100, nz=70, n_rad_in=20,
               n_lambda=50, n_T=100, dark_zone=True, physical_length=None, isotropic=False):
    """G1: ref4.1.para geometry (cylindrical 100x70x1, disk 1-300 AU, H=10 AU at
    100 AU, beta=1.125, p=-0.5, star 5000 K / 2 Rsun blackbody, 50 wavelengths
    0.1-3000 um, n_T=100) with the synthetic optics."""
    zones = [DiskZone()]
    P = cylindrical_grid(n_rad, nz, 1, n_rad_in, zones, l3D=False)
    _finish(P, disk_density(P, zones), n_lambda, n_T, tau_mid, pola, isotropic, n_photons_eq_th, 5000.0, 2.0)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    if dark_zone:
        P.l_dark_zone = define_dark_zone(P, P.lambda_seuil, 1500.0, physical_length)
    repartition_energie(P)
    P.name = "ref4.1-like (G1)"
    return P


def pascucci_optics(lam, pola=True, isotropic=True):
    """Stand-in for the single 0.12 um astronomical-silicate grain of the Pascucci et al. (2004) benchmark (the
    reference computes it with Mie theory on Draine_Si.dat, which is not available offline): geometric cross sections
    below 2 pi a = 0.754 um, absorption ~ 1/lambda beyond with the 9.7 and 18 um silicate bands and ~ 1/lambda^2 in
    the far infrared, Rayleigh scattering ~ 1/lambda^4, isotropic phase function (benchmarks.f90:30-31 forces
    lisotropic).  Same return convention as synthetic_optics (unit extinction at the shortest wavelengths)."""
    n = len(lam)
    xs = 0.754 / np.maximum(lam, 1e-30)
    q_abs = np.minimum(1.0, xs) * (1.0 + 2.5 * np.exp(-(((lam - 9.7) / 1.5) ** 2)) + 1.0 * np.exp(-(((lam - 18.0) / 4.0) ** 2)))
    q_abs = q_abs * np.where(lam > 30.0, 30.0 / lam, 1.0)
    q_sca = np.minimum(1.0, xs ** 4)
    kext = (q_abs + q_sca) / 2.0
    albedo = q_sca / (q_abs + q_sca)
    g = 0.0 * lam
    s11 = np.ones((NANG_SCATT + 1, n))
    theta = np.arange(NANG_SCATT + 1) * PI / NANG_SCATT
    mu = np.cos(theta)
    pol = (-(1.0 - mu * mu) / (1.0 + mu * mu), np.ones_like(mu), 2.0 * mu / (1.0 + mu * mu), 0.0 * mu, 2.0 * mu / (1.0 + mu * mu)) if pola else None
    return kext, albedo, g, s11, pol


def pascucci_like(tau_V=1.0, n_photons_eq_th=10000, n_rad=100, nz=70, n_rad_in=20, n_T=100, pola=True):
    """G2: Pascucci_3.0.para -- the 2D disk benchmark of Pascucci et al. 2004: cylindrical 100x70x1, disk 1-1000 AU,
    scale height 99.74 AU at 500 AU, flaring 1.125, surface-density exponent +0.125 (rho ~ 1/r), one 0.12 um
    silicate grain with isotropic scattering, 61 wavelengths 0.1107-2168.8 um, star 5800 K / 1 Rsun, 1.28e6 thermal
    packets (Pascucci_3.0.para:4-8,49-56,68; benchmarks.f90:30-31).  tau_V is the midplane optical depth at
    0.55 um from the star to the outer edge (0.1, 1, 10, 100 in the benchmark)."""
    zones = [DiskZone(rin=1.0, rout=1000.0, sclht=99.73557010035817, rref=500.0, exp_beta=1.125, surf=0.125, dust_mass=1.1e-6)]
    P = cylindrical_grid(n_rad, nz, 1, n_rad_in, zones, l3D=False)
    _finish(P, disk_density(P, zones), 61, n_T, tau_V, pola, True, n_photons_eq_th, 5800.0, 1.0,
            lam_range=(0.110662, 2168.76), lam_tau=0.55, optics=pascucci_optics)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    repartition_energie(P)
    P.name = "Pascucci-like (G2), tau_V = %g" % tau_V
    return P


def ref41_3d_like(n_photons_eq_th=1000, tau_mid=1.0e3, n_rad=100, nz=50, n_az=72, n_rad_in=20,
                  n_lambda=50, n_T=100, pola=False, spiral=0.5):
    """G4: ref4.1_3D.para geometry: 100 x (2x50) x 72 = 720 000 cells, no dark zone
    (dust_transfer.f90:290-293)."""
    zones = [DiskZone()]
    P = cylindrical_grid(n_rad, nz, n_az, n_rad_in, zones, l3D=True)
    rho = disk_density(P, zones)
    if spiral:      # m=2 spiral perturbation so that azimuthal walls matter
        rho = rho * (1.0 + spiral * np.cos(2.0 * (P.phi_grid - np.log(P.r_grid))))
    _finish(P, rho, n_lambda, n_T, tau_mid, pola, False, n_photons_eq_th, 5000.0, 2.0)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    repartition_energie(P)
    P.name = "ref4.1_3D-like (G4)"
    return P


def spherical_shell(n_photons_eq_th=1000, tau_mid=10.0, n_rad=60, nz=30, n_az=1, l3D=False,
                    n_lambda=50, n_T=100, pola=False, isotropic=False):
    """debris.para-style spherical grid (test_data/debris/debris.para:15)."""
    P = spherical_grid(n_rad, nz, n_az, 5, 10.0, 200.0, l3D=l3D)
    _finish(P, envelope_density(P, 0.5 if l3D else 0.0), n_lambda, n_T, tau_mid, pola, isotropic, n_photons_eq_th, 5000.0, 2.0)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    repartition_energie(P)
    P.name = "spherical shell"
    return P


# ---------------------------------------------------------------------------
# caller-side post-processing (stays Fortran in production; numpy here for tests)
# ---------------------------------------------------------------------------
def temp_finale(P, xKJ_abs, n_ranks_total_scale=1.0):
    """Temp_finale + Temp_LTE(id=0), thermal_emission.f90:649-706,870-906:
    final dust temperature from the merged absorbed-energy tally."""
    xKJ = np.asarray(xKJ_abs, np.float64) * n_ranks_total_scale
    Qheat = xKJ * P.L_packet_th / P.volume
    logQ = P.log_Qcool_minus_extra_heating
    pnc = P.p_n_cells
    T = np.full(P.n_cells, float(P.T_min))
    ltab = np.log(P.tab_Temp.astype(np.float64))
    ok = Qheat >= np.finfo(np.float64).tiny
    lq = np.log(np.where(ok, Qheat, 1.0))
    for ic in np.nonzero(ok)[0]:
        col = logQ[:, ic if pnc > 1 else 0]
        if lq[ic] < col[0]:
            continue
        Ti = int(np.searchsorted(col, lq[ic], side="left")) + 1      # first Ti with col(Ti) >= lq, 1-based
        Ti = min(max(Ti, 2), P.n_T)
        frac = (lq[ic] - col[Ti - 2]) / (col[Ti - 1] - col[Ti - 2])
        T[ic] = np.exp(ltab[Ti - 1] * frac + ltab[Ti - 2] * (1.0 - frac))
    return T


# ---------------------------------------------------------------------------
# Voronoi mesh (stand-in for voro++: the reference's tessellation library is not vendored)
# ---------------------------------------------------------------------------
def voronoi_mesh(n_points=1500, half_size=100.0, seed=12345, cut_fraction=0.1):
    """Seeds drawn from a disk-like density inside a box, tessellated with scipy (qhull).

    The six walls are produced exactly by the mirror trick: every seed is reflected across each wall,
    so the bisector between a seed and its image IS the wall and every real cell is bounded and inside
    the box.  Neighbour lists follow the reference's layout (Voronoi.f90:23-66): 1-based cell ids,
    -wall_id for walls, first/last pointers into `neighbours_list`.  A fraction of elongated cells is
    flagged `was_cut` (the reference cuts cells whose faces are farther than 3h, voro++_wrapper.cpp:209-225)
    so that the cut-sphere branch of cross_Voronoi_cell is exercised."""
    from scipy.spatial import Voronoi, ConvexHull
    rng = np.random.default_rng(seed)
    L = half_size
    # disk-like seeds: gaussian in z with flaring, r^-1 in radius; keep away from the walls
    r = L * 0.9 * rng.uniform(0.05, 1.0, n_points)
    phi = rng.uniform(0, 2 * np.pi, n_points)
    z = rng.normal(0.0, 0.15 * r)
    pts = np.stack([r * np.cos(phi), r * np.sin(phi), np.clip(z, -0.9 * L, 0.9 * L)], axis=1)
    pts = np.clip(pts, -0.95 * L, 0.95 * L)
    n = n_points
    lim = np.array([-L, L, -L, L, -L, L])
    ext = [pts]
    for wdx in range(6):
        ax, val = wdx // 2, lim[wdx]
        m = pts.copy()
        m[:, ax] = 2.0 * val - m[:, ax]
        ext.append(m)
    allp = np.concatenate(ext)
    vor = Voronoi(allp)
    neigh = [set() for _ in range(n)]
    for a, b in vor.ridge_points:
        for i, j in ((a, b), (b, a)):
            if i < n:
                neigh[i].add(int(j) + 1 if j < n else -(int(j - n) // n + 1))
    first = np.zeros(n, np.int32); last = np.zeros(n, np.int32)
    flat = []
    for i in range(n):
        lst = sorted(neigh[i], key=lambda q: (q < 0, abs(q)))
        first[i] = len(flat) + 1
        flat.extend(lst)
        last[i] = len(flat)
    vol = np.zeros(n); h = np.zeros(n); elong = np.zeros(n)
    for i in range(n):
        verts = vor.vertices[vor.regions[vor.point_region[i]]]
        vol[i] = ConvexHull(verts).volume
        h[i] = (3.0 * vol[i] / (4.0 * np.pi)) ** (1.0 / 3.0)
        elong[i] = np.max(np.linalg.norm(verts - pts[i], axis=1)) / h[i]
    P = Problem()
    P.kind, P.l3D = MCB_GRID_VORONOI, 1
    P.n_rad, P.nz, P.n_az = 0, 0, 0
    P.n_cells = n
    P.Rmax2, P.zmaxmax = 3.0 * L * L, L
    P.vor_xyz = np.asfortranarray(pts.T.copy())        # (3, n)
    P.vor_h = h
    P.vor_first, P.vor_last = first, last
    P.neighbours_list = np.array(flat, np.int32)
    thr = np.quantile(elong, 1.0 - cut_fraction) if cut_fraction > 0 else np.inf
    P.vor_was_cut = (elong > thr).astype(np.int32)
    P.vor_is_star = np.zeros(n, np.int32)
    P.vor_is_star_neighbour = np.zeros(n, np.int32)
    P.wall_x = [[-1, 0, 0, -L], [1, 0, 0, L], [0, -1, 0, -L], [0, 1, 0, L], [0, 0, -1, -L], [0, 0, 1, L]]
    P.cutting_distance_o_h = 1.6
    P.volume = vol
    P.r_grid = np.hypot(pts[:, 0], pts[:, 1]); P.z_grid = pts[:, 2]; P.phi_grid = np.arctan2(pts[:, 1], pts[:, 0])
    P.r_lim = P.r_lim_2 = P.r_lim_3 = P.z_lim = P.zmax = P.tan_theta_lim = P.theta_lim = P.tan_phi_lim = None
    P.cell_map_i = P.cell_map_j = P.cell_map_k = None
    P.n_cells_tot = 0
    return P


def voronoi_disk(n_points=1500, n_photons_eq_th=200, tau_mid=30.0, n_lambda=50, n_T=100, pola=False, seed=12345):
    """A small disk on a Voronoi mesh (the phantom / SPH use case, test size)."""
    P = voronoi_mesh(n_points, seed=seed)
    rs = np.maximum(P.r_grid, 1.0)
    rho = rs ** -1.5 * np.exp(-0.5 * (P.z_grid / (0.15 * rs)) ** 2) + 1e-6
    return _finish_voronoi(P, rho, n_photons_eq_th, tau_mid, n_lambda, n_T, pola, "Voronoi disk")


def _finish_voronoi(P, rho, n_photons_eq_th, tau_mid, n_lambda, n_T, pola, name):
    """star, constant-dust opacities (p_n_cells = 1) and emission tables of a Voronoi model with cell densities rho"""
    # _finish needs a radial column to scale kappa: use a pseudo column through the cells sorted by radius
    order = np.argsort(P.r_grid)
    P.n_rad = 0
    P.n_lambda, P.n_T = n_lambda, n_T
    P.tab_lambda, P.tab_delta_lambda = init_lambda(n_lambda)
    P.tab_lambda = np.float32(P.tab_lambda).astype(np.float64)
    P.tab_delta_lambda = np.float32(P.tab_delta_lambda).astype(np.float64)
    P.T_min, P.T_max = 1.0, 3000.0
    P.tab_Temp = init_tab_Temp(n_T, P.T_min, P.T_max)
    P.n_photons_loop, P.n_photons_eq_th = 128, n_photons_eq_th
    P.n_stars = 1
    P.star_xyzr = np.asfortranarray(np.array([[0.0], [0.0], [0.0], [2.0 * RSUN_TO_AU]]))
    P.star_T = np.array([5000.0])
    P.star_out_model = np.zeros(1, np.int32)
    d2 = np.sum(P.vor_xyz ** 2, axis=0)
    P.star_icell = np.array([int(np.argmin(d2)) + 1], np.int32)       # the cell that contains the star
    P.p_n_cells, P.p_n_lambda_pos = 1, n_lambda
    P.kappa_factor = rho / rho.max()
    kext, albedo, g, s11, pol = synthetic_optics(P.tab_lambda, pola=pola)
    l_seuil = int(np.argmax(P.tab_lambda > 0.81)) + 1
    P.lambda_seuil = l_seuil
    rr = P.r_grid[order]
    col = float(np.sum(P.kappa_factor[order][:-1] * np.diff(rr)))
    k0 = tau_mid / (col * kext[l_seuil - 1])
    P.kappa = np.asfortranarray((k0 * kext).reshape(1, n_lambda))
    P.tab_albedo_pos = np.asfortranarray(np.float32(albedo).reshape(1, n_lambda))
    P.tab_g_pos = np.asfortranarray(np.float32(g).reshape(1, n_lambda))
    P.kappa_abs_LTE = np.asfortranarray(P.kappa * (1.0 - P.tab_albedo_pos.astype(np.float64)))
    for k, v in scattering_tables(P, s11, albedo, P.kappa[0], pol).items():
        setattr(P, k, v)
    init_reemission(P)
    star_energy(P)
    P.l_dark_zone = np.zeros(P.n_cells, np.int32)
    P.E_paquet = 1.0
    P.R_ISM = 0.0
    P.centre_ISM = (0.0, 0.0, 0.0)
    repartition_energie(P)
    P.name = name
    return P


_MESH_KEYS = ("vor_xyz", "vor_h", "vor_first", "vor_last", "neighbours_list", "vor_was_cut", "volume", "wall_x", "rho", "Rmax2", "zmaxmax")


def voronoi_sph_disk(n_points=1000000, n_photons_eq_th=1000, tau_mid=1.0e3, n_lambda=50, n_T=100, pola=False, seed=12345,
                     keep=0.999, zone=None, cache=None):
    """G5: a Voronoi mesh on the particles of a synthetic SPH disk (phantom-like), SURVEY 8d.

    Particles: n_points drawn from the G1 density (ref4.1 geometry: Sigma ~ r^-0.5 between 1 and 300 AU, Gaussian in z
    with H = 10 AU (r / 100 AU)^1.125), numpy seed 12345 (inverse-transform sampling of the same distribution the
    survey's rejection sampling would give).  Box: the reference's percentile rule, SPH_keep_particles = 0.999, i.e. the
    0.05 % outermost particles are dropped on each side of each axis (SPH2mcfost.f90:246,261-273).  Mesh: scipy Delaunay
    (qhull) gives the neighbour lists voro++ gives the reference (Voronoi.f90:472-475); cells on the hull get the walls
    they face as negative neighbours.  Stand-ins, documented: the cell volume is the expectation 1 / (N pdf(x_i)) of the
    Voronoi volume under the sampling density instead of the polyhedron's (a million convex hulls in Python), h =
    (3 V / 4 pi)^(1/3), and a cell is `was_cut` when its farthest neighbour is more than 6 h away (the reference cuts
    faces farther than 3 h, voro++_wrapper.cpp:209-225)."""
    import os
    if cache and os.path.exists(cache):      # a mesh generated earlier with the same arguments (tools/make_g5_mesh.py)
        d = np.load(cache)
        P = Problem()
        P.kind, P.l3D = MCB_GRID_VORONOI, 1
        P.n_rad, P.nz, P.n_az = 0, 0, 0
        P.vor_xyz = np.asfortranarray(d["vor_xyz"]); P.n_cells = P.vor_xyz.shape[1]
        for k in ("vor_h", "vor_first", "vor_last", "neighbours_list", "vor_was_cut", "volume"):
            setattr(P, k, d[k])
        P.wall_x = [list(map(float, w)) for w in d["wall_x"]]
        P.Rmax2, P.zmaxmax = float(d["Rmax2"]), float(d["zmaxmax"])
        P.vor_is_star = np.zeros(P.n_cells, np.int32); P.vor_is_star_neighbour = np.zeros(P.n_cells, np.int32)
        P.cutting_distance_o_h = 3.0
        P.r_grid = np.hypot(P.vor_xyz[0], P.vor_xyz[1]); P.z_grid = P.vor_xyz[2].copy(); P.phi_grid = np.arctan2(P.vor_xyz[1], P.vor_xyz[0])
        P.r_lim = P.r_lim_2 = P.r_lim_3 = P.z_lim = P.zmax = P.tan_theta_lim = P.theta_lim = P.tan_phi_lim = None
        P.cell_map_i = P.cell_map_j = P.cell_map_k = None
        P.n_cells_tot = 0
        return _finish_voronoi(P, d["rho"], n_photons_eq_th, tau_mid, n_lambda, n_T, pola, "Voronoi SPH disk (G5), %d cells" % P.n_cells)
    from scipy.spatial import Delaunay
    z_ = zone or DiskZone()
    rng = np.random.default_rng(seed)
    # r from p(r) ~ r Sigma(r) = r^(1+surf) on [rin, rout]; z Gaussian with H(r); phi uniform
    a = 2.0 + z_.surf
    uu = rng.uniform(size=n_points)
    r = (z_.rin ** a + uu * (z_.rout ** a - z_.rin ** a)) ** (1.0 / a)
    H = z_.sclht * (r / z_.rref) ** z_.exp_beta
    zz = rng.normal(size=n_points) * H
    phi = rng.uniform(0.0, 2.0 * np.pi, n_points)
    pts = np.stack([r * np.cos(phi), r * np.sin(phi), zz], axis=1)
    # percentile box (SPH2mcfost.f90:261-273)
    lo = np.quantile(pts, 0.5 * (1.0 - keep), axis=0)
    hi = np.quantile(pts, 1.0 - 0.5 * (1.0 - keep), axis=0)
    ok = np.all((pts > lo) & (pts < hi), axis=1)
    pts, r, H, zz = pts[ok], r[ok], H[ok], zz[ok]
    n = len(pts)
    tri = Delaunay(pts)
    indptr, indices = tri.vertex_neighbor_vertices
    deg = np.diff(indptr).astype(np.int64)
    # walls faced by the hull cells: a hull vertex gets every wall it is closer to than to its farthest neighbour
    hull = np.unique(tri.convex_hull)
    dmax = np.zeros(n)
    src = np.repeat(np.arange(n), deg)
    dist = np.sqrt(np.sum((pts[src] - pts[indices]) ** 2, axis=1))
    np.maximum.at(dmax, src, dist)
    wall_lists = [[] for _ in range(6)]
    dw = np.stack([pts[hull, 0] - lo[0], hi[0] - pts[hull, 0], pts[hull, 1] - lo[1], hi[1] - pts[hull, 1], pts[hull, 2] - lo[2], hi[2] - pts[hull, 2]], axis=1)
    near = dw <= np.maximum(dmax[hull][:, None], dw.min(axis=1, keepdims=True))
    n_wall = np.zeros(n, np.int64)
    n_wall[hull] = near.sum(axis=1)
    first = np.zeros(n, np.int64)
    tot = deg + n_wall
    first[1:] = np.cumsum(tot)[:-1]
    flat = np.zeros(int(tot.sum()), np.int32)
    # cell neighbours first (1-based ids), then the walls (negative ids)
    pos = first[src] + (np.arange(len(src)) - indptr[src])
    flat[pos] = indices.astype(np.int32) + 1
    hrow, hwall = np.nonzero(near)
    order = np.lexsort((hwall, hrow))
    hrow, hwall = hrow[order], hwall[order]
    rank = np.arange(len(hrow)) - np.searchsorted(hrow, hrow, side="left")
    flat[first[hull[hrow]] + deg[hull[hrow]] + rank] = -(hwall.astype(np.int32) + 1)
    # expected Voronoi volume under the sampling density
    norm_r = (z_.rout ** a - z_.rin ** a) / a
    pdf = (r ** (a - 1.0) / norm_r) / (2.0 * np.pi * r) * np.exp(-0.5 * (zz / H) ** 2) / (np.sqrt(2.0 * np.pi) * H)
    vol = 1.0 / (n_points * pdf)
    h = (3.0 * vol / (4.0 * np.pi)) ** (1.0 / 3.0)
    P = Problem()
    P.kind, P.l3D = MCB_GRID_VORONOI, 1
    P.n_rad, P.nz, P.n_az = 0, 0, 0
    P.n_cells = n
    L = float(np.max(np.abs(np.concatenate([lo, hi]))))
    P.Rmax2, P.zmaxmax = 3.0 * L * L, float(max(abs(lo[2]), abs(hi[2])))
    P.vor_xyz = np.asfortranarray(pts.T.copy())
    P.vor_h = h
    P.vor_first = (first + 1).astype(np.int32); P.vor_last = (first + tot).astype(np.int32)
    P.neighbours_list = flat
    P.vor_was_cut = (dmax > 6.0 * h).astype(np.int32)
    P.vor_is_star = np.zeros(n, np.int32)
    P.vor_is_star_neighbour = np.zeros(n, np.int32)
    P.wall_x = [[-1, 0, 0, lo[0]], [1, 0, 0, hi[0]], [0, -1, 0, lo[1]], [0, 1, 0, hi[1]], [0, 0, -1, lo[2]], [0, 0, 1, hi[2]]]
    P.cutting_distance_o_h = 3.0
    P.volume = vol
    P.r_grid = np.hypot(pts[:, 0], pts[:, 1]); P.z_grid = pts[:, 2]; P.phi_grid = np.arctan2(pts[:, 1], pts[:, 0])
    P.r_lim = P.r_lim_2 = P.r_lim_3 = P.z_lim = P.zmax = P.tan_theta_lim = P.theta_lim = P.tan_phi_lim = None
    P.cell_map_i = P.cell_map_j = P.cell_map_k = None
    P.n_cells_tot = 0
    rho = pdf / pdf.max()          # the particles ARE the density: equal-mass particles, rho_i = m / V_i
    if cache:
        os.makedirs(os.path.dirname(os.path.abspath(cache)), exist_ok=True)
        np.savez(cache, vor_xyz=P.vor_xyz, vor_h=np.float64(h), vor_first=P.vor_first, vor_last=P.vor_last, neighbours_list=flat,
                 vor_was_cut=P.vor_was_cut, volume=vol, wall_x=np.array(P.wall_x, np.float64), rho=rho, Rmax2=P.Rmax2, zmaxmax=P.zmaxmax)
    return _finish_voronoi(P, rho, n_photons_eq_th, tau_mid, n_lambda, n_T, pola, "Voronoi SPH disk (G5), %d cells" % n)


def ref41_multi_like(n_photons_eq_th=200, n_rad=40, nz=20, n_rad_in=5, tau_mid=300.0, n_lambda=50, n_T=100):
    """G3-like (ref4.1_multi.para, LTE part): two zones (1-5 AU, 10-300 AU) with DIFFERENT dust, so
    lvariable_dust = .true. and every opacity / scattering / thermal table is per cell
    (p_n_cells = n_cells, kappa_factor = 1, init_mcfost.f90:1582, dust_prop.f90:951-955).  The
    scattering tables use the single-wavelength layout p_n_lambda_pos = 1 that the reference falls back
    to when per-cell x per-wavelength tables would not fit (scattering.f90:39-66)."""
    zones = [DiskZone(rin=1.0, rout=5.0, dust_mass=1e-5), DiskZone(rin=10.0, rout=300.0, dust_mass=1e-3)]
    P = cylindrical_grid(n_rad, nz, 1, n_rad_in, zones, l3D=False)
    P.n_lambda, P.n_T = n_lambda, n_T
    P.tab_lambda, P.tab_delta_lambda = init_lambda(n_lambda)
    P.tab_lambda = np.float32(P.tab_lambda).astype(np.float64)
    P.tab_delta_lambda = np.float32(P.tab_delta_lambda).astype(np.float64)
    P.T_min, P.T_max = 1.0, 3000.0
    P.tab_Temp = init_tab_Temp(n_T, P.T_min, P.T_max)
    P.n_photons_loop, P.n_photons_eq_th = 128, n_photons_eq_th
    P.n_stars = 1
    P.star_xyzr = np.asfortranarray(np.array([[0.0], [0.0], [0.0], [2.0 * RSUN_TO_AU]]))
    P.star_T = np.array([5000.0]); P.star_out_model = np.zeros(1, np.int32)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    nc = P.n_cells
    rho = disk_density(P, zones)
    inner = P.r_grid < 7.0
    kext, albedo, g, s11, _ = synthetic_optics(P.tab_lambda, pola=False)
    # zone A (inner): small grains -> steeper opacity law, higher albedo; zone B: the G1 optics
    kextA = np.where(P.tab_lambda > 0.3, (0.3 / P.tab_lambda) ** 1.6, 1.0)
    albA = np.clip(albedo * 1.6, 0.0, 0.9)
    gA = g * 0.3
    l_seuil = int(np.argmax(P.tab_lambda > 0.81)) + 1
    P.lambda_seuil = l_seuil
    mid = np.arange(n_rad)
    col = float(np.sum((rho / rho.max())[mid] * (P.r_lim[1:] - P.r_lim[:-1])))
    k0 = tau_mid / (col * kext[l_seuil - 1])
    dens = rho / rho.max()
    P.p_n_cells, P.p_n_lambda_pos = nc, 1
    P.kappa_factor = np.ones(nc)
    kap = np.where(inner[:, None], kextA[None, :], kext[None, :]) * dens[:, None] * k0
    alb = np.where(inner[:, None], albA[None, :], albedo[None, :])
    gg = np.where(inner[:, None], gA[None, :], g[None, :])
    P.kappa = np.asfortranarray(kap)
    P.tab_albedo_pos = np.asfortranarray(np.float32(alb))
    P.tab_g_pos = np.asfortranarray(np.float32(gg))
    P.kappa_abs_LTE = np.asfortranarray(kap * (1.0 - alb))
    # per-cell s11 CDF at the (single) tabulated wavelength index 1
    theta = np.arange(NANG_SCATT + 1) * PI / NANG_SCATT
    prob = np.zeros((NANG_SCATT + 1, nc, 1), np.float32, order="F")
    tab = np.zeros((NANG_SCATT + 1, nc, 1), np.float32, order="F")
    for zone_mask, gz in ((inner, gA[0]), (~inner, g[0])):
        s = (1.0 - gz ** 2) * (1.0 + gz ** 2 - 2.0 * gz * np.cos(theta)) ** (-1.5)
        w = s * np.sin(theta)
        c = np.concatenate(([0.0], np.cumsum(0.5 * (w[1:] + w[:-1]))))
        c = c / c[-1]
        prob[:, zone_mask, 0] = np.float32(c)[:, None]
        tab[:, zone_mask, 0] = np.float32(s / (np.sum(w) * 2.0 * PI))[:, None]
    P.prob_s11_pos, P.tab_s11_pos = prob, tab
    for nm in ("tab_s12_o_s11_pos", "tab_s22_o_s11_pos", "tab_s33_o_s11_pos", "tab_s34_o_s11_pos", "tab_s44_o_s11_pos"):
        setattr(P, nm, None)
    init_reemission(P)
    star_energy(P)
    P.l_dark_zone = np.zeros(nc, np.int32)
    P.E_paquet = 1.0; P.R_ISM = 0.0; P.centre_ISM = (0.0, 0.0, 0.0)
    repartition_energie(P)
    P.name = "ref4.1_multi-like (G3, LTE part)"
    return P


AU_TO_CM = 149597870700.0 * 100.0
MUM_TO_CM = 1.0e-4
AU_TO_CM_MUM2 = AU_TO_CM * (MUM_TO_CM * MUM_TO_CM)      # AU_to_cm * mum_to_cm**2


def multi_grain_like(n_photons_eq_th=100, n_rad=24, nz=12, n_rad_in=4, tau_mid=30.0, n_lambda=30, n_T=60,
                     n_LTE=5, n_nLTE=3, n_nRE=3, variable=False, pola=True, qre_fraction=0.5, seed=11):
    """A disk whose dust is a SIZE DISTRIBUTION of n_LTE + n_nLTE + n_nRE grains in the three heating
    regimes of ref4.1_multi.para (methode_chauffage 1, 2, 3): per-grain cross sections, phase functions
    and emissivities, and every table the per-grain branches read, built the way the reference builds
    them (opacite dust_prop.f90:787-1000, init_reemission thermal_emission.f90:404-644,
    update_proba_abs_nRE :1518-1600).  Optics are an analytic stand-in for Mie theory:
    Q_abs = x/(1+x), Q_sca = x^4/(1+x^4), g = 0.7 x^2/(1+x^2) with x = 2 pi a / lambda.
    variable=True: lvariable_dust (size-dependent settling, p_n_cells = n_cells, p_n_lambda_pos = 1)."""
    zones = [DiskZone(rin=1.0, rout=100.0, edge=0.0)]
    P = cylindrical_grid(n_rad, nz, 1, n_rad_in, zones, l3D=False)
    nc = P.n_cells
    P.n_lambda, P.n_T = n_lambda, n_T
    P.tab_lambda, P.tab_delta_lambda = init_lambda(n_lambda)
    P.tab_lambda = np.float32(P.tab_lambda).astype(np.float64)
    P.tab_delta_lambda = np.float32(P.tab_delta_lambda).astype(np.float64)
    P.T_min, P.T_max = 1.0, 3000.0
    P.tab_Temp = init_tab_Temp(n_T, P.T_min, P.T_max)
    P.n_photons_loop, P.n_photons_eq_th = 128, n_photons_eq_th
    P.n_stars = 1
    P.star_xyzr = np.asfortranarray(np.array([[0.0], [0.0], [0.0], [2.0 * RSUN_TO_AU]]))
    P.star_T = np.array([6000.0]); P.star_out_model = np.zeros(1, np.int32)
    P.star_icell = np.array([star_icell_analytic(P)], np.int32)
    lam = P.tab_lambda
    # ---- grains: big (LTE) first, then nLTE, then the smallest (nRE), like the reference's sorted populations
    K = n_LTE + n_nLTE + n_nRE
    a = np.logspace(1.0, -2.3, K)                       # um
    ng = a ** -3.5 * a                                   # dn/dlog a
    ng = ng / ng.sum()
    P.n_grains_tot, P.n_grains = K, ng
    P.grain_RE_LTE_start, P.grain_RE_LTE_end = 1, n_LTE
    P.grain_RE_nLTE_start, P.grain_RE_nLTE_end = n_LTE + 1, n_LTE + n_nLTE
    P.grain_nRE_start, P.grain_nRE_end = n_LTE + n_nLTE + 1, K
    P.grain_zone = np.ones(K, np.int32)
    x = 2.0 * PI * a[:, None] / lam[None, :]
    geo = PI * a[:, None] ** 2
    C_abs = np.float32(geo * x / (1.0 + x))
    C_sca = np.float32(geo * x ** 4 / (1.0 + x ** 4))
    C_ext = np.float32(C_abs.astype(np.float64) + C_sca)
    tab_g = np.float32(0.7 * x ** 2 / (1.0 + x ** 2))
    P.C_abs, P.C_sca, P.tab_g = (np.asfortranarray(t) for t in (C_abs, C_sca, tab_g))
    P.C_abs_norm = np.asfortranarray(np.float32(C_abs.astype(np.float64) * AU_TO_CM * MUM_TO_CM ** 2))
    # ---- number densities: dust_density_o_n_grains(n_dens, n_cells)
    rho = disk_density(P, zones)
    rho = rho / rho[0]
    if variable:
        # settling: scale height shrinks with grain size
        h_fac = np.clip((a / a.min()) ** -0.15, 0.3, 1.0)
        z_o_h = np.abs(P.z_grid) / np.maximum(zones[0].sclht * (P.r_grid / zones[0].rref) ** zones[0].exp_beta, 1e-30)
        dd = rho[None, :] * np.exp(-0.5 * z_o_h[None, :] ** 2 * (1.0 / h_fac[:, None] ** 2 - 1.0)) / h_fac[:, None]
        P.n_dens = K
    else:
        dd = rho[None, :].copy()
        P.n_dens = 1
    pk = (np.arange(K) if variable else np.zeros(K, int))
    dens = dd[pk, :] * ng[:, None]                       # (K, n_cells): density of grain k in each cell
    # ---- opacities (dust_prop.f90:850-890, x fact at :965-970), scaled to tau_mid at 0.81 um
    l_seuil = int(np.argmax(lam > 0.81)) + 1
    P.lambda_seuil = l_seuil
    kext_cell = np.einsum("kl,kc->cl", C_ext.astype(np.float64), dens)          # (n_cells, n_lambda), um^2 cm^-3
    mid = np.arange(n_rad)
    col = float(np.sum(kext_cell[mid, l_seuil - 1] * (P.r_lim[1:] - P.r_lim[:-1])))
    scale = tau_mid / (col * AU_TO_CM_MUM2)
    dd = dd * scale; dens = dens * scale
    P.dust_density_o_n_grains = np.asfortranarray(dd)

    def ksum(C, ks, cells):
        return np.einsum("kl,kc->cl", C[ks].astype(np.float64), dens[ks][:, cells])
    cells = np.arange(nc) if variable else np.array([0])
    allk = np.arange(K); kL = np.arange(0, n_LTE); kN = np.arange(n_LTE, n_LTE + n_nLTE); kR = np.arange(n_LTE + n_nLTE, K)
    P.p_n_cells = nc if variable else 1
    P.kappa_factor = np.ones(nc) if variable else rho.copy()
    kap = ksum(C_ext, allk, cells); ksca = ksum(C_sca, allk, cells)
    P.kappa = np.asfortranarray(kap * AU_TO_CM_MUM2)
    alb = np.where(kap > 0, ksca / np.maximum(kap, 1e-300), 0.0)
    P.tab_albedo_pos = np.asfortranarray(np.float32(alb))
    gpos = np.einsum("kl,kc->cl", (C_sca * tab_g).astype(np.float64), dens[:, cells]) / np.maximum(ksca, 1e-300)
    P.tab_g_pos = np.asfortranarray(np.float32(gpos))
    P.kappa_abs_LTE = np.asfortranarray(ksum(C_abs, kL, cells) * AU_TO_CM_MUM2)
    P.kappa_abs_nLTE = np.asfortranarray(ksum(C_abs, kN, cells) * AU_TO_CM_MUM2)
    # ---- per-grain phase functions (HG per grain; tab_s11 = 1 after mueller's normalisation)
    theta = np.arange(NANG_SCATT + 1) * PI / NANG_SCATT
    mu = np.cos(theta)
    gk = tab_g.astype(np.float64)
    s11 = (1.0 - gk[None] ** 2) * (1.0 + gk[None] ** 2 - 2.0 * gk[None] * mu[:, None, None]) ** (-1.5)      # (181, K, nl)
    wgt = s11 * np.sin(theta)[:, None, None]
    cdf = np.concatenate((np.zeros((1, K, n_lambda)), np.cumsum(0.5 * (wgt[1:] + wgt[:-1]), axis=0)))
    cdf = cdf / cdf[-1:]
    P.prob_s11 = np.asfortranarray(np.float32(np.transpose(cdf, (2, 1, 0))))    # (n_lambda, K, 0:180)
    ones = np.ones((NANG_SCATT + 1, K, n_lambda), np.float32)
    P.tab_s11 = np.asfortranarray(ones)
    if pola:
        pmax = 0.2 + 0.3 / (1.0 + x)                      # small grains polarise more
        P.tab_s12 = np.asfortranarray(np.float32(-pmax[None] * ((1.0 - mu * mu) / (1.0 + mu * mu))[:, None, None]))
        P.tab_s22 = np.asfortranarray(ones.copy())
        s33 = (2.0 * mu / (1.0 + mu * mu))[:, None, None] * np.ones((1, K, n_lambda))
        P.tab_s33 = np.asfortranarray(np.float32(s33)); P.tab_s44 = np.asfortranarray(np.float32(s33))
        P.tab_s34 = np.asfortranarray(np.float32((0.1 * np.sin(theta) ** 2 * mu)[:, None, None] * np.ones((1, K, n_lambda))))
    # ksca_CDF(0:K, p_n_cells, n_lambda) (dust_prop.f90:1180-1200): normalised cumulative of C_sca * density
    c = np.cumsum(np.einsum("kl,kc->kcl", C_sca.astype(np.float64), dens[:, cells]), axis=0)
    c = np.concatenate((np.zeros((1,) + c.shape[1:]), c)) / np.maximum(c[-1:], 1e-300)
    P.ksca_CDF = np.asfortranarray(c)
    # ---- method-2 tables of the population (calc_local_scattering_matrices)
    if variable:
        P.p_n_lambda_pos = 1
        w = (C_sca.astype(np.float64)[:, None, 0] * dens)                        # (K, n_cells) at lambda index 1
        s_cell = np.einsum("akl,kc->ac", s11[:, :, :1], w) / np.maximum(w.sum(0), 1e-300)
        wg = s_cell * np.sin(theta)[:, None]
        cc = np.concatenate((np.zeros((1, nc)), np.cumsum(0.5 * (wg[1:] + wg[:-1]), axis=0)))
        P.prob_s11_pos = np.asfortranarray(np.float32(cc / cc[-1:]).reshape(NANG_SCATT + 1, nc, 1))
        P.tab_s11_pos = np.asfortranarray(np.float32(s_cell / (wg.sum(0) * 2.0 * PI)).reshape(NANG_SCATT + 1, nc, 1))
        for nm in ("tab_s12_o_s11_pos", "tab_s22_o_s11_pos", "tab_s33_o_s11_pos", "tab_s34_o_s11_pos", "tab_s44_o_s11_pos"):
            setattr(P, nm, None)
    else:
        P.p_n_lambda_pos = n_lambda
        w = C_sca.astype(np.float64) * dens[:, :1]                               # (K, nl)
        s_pop = np.einsum("akl,kl->al", s11, w) / np.maximum(w.sum(0), 1e-300)
        pol = None
        if pola:
            pol = ((-0.3 * (1.0 - mu * mu) / (1.0 + mu * mu)), np.ones_like(mu), 2.0 * mu / (1.0 + mu * mu),
                   0.1 * np.sin(theta) ** 2 * mu, 2.0 * mu / (1.0 + mu * mu))
        for k_, v_ in scattering_tables(P, s_pop, alb[0], P.kappa[0], pol).items():
            setattr(P, k_, v_)
    # ---- LTE thermal tables of the LTE grains, emission tables
    init_reemission(P)
    star_energy(P)
    P.l_dark_zone = np.zeros(nc, np.int32)
    P.E_paquet = 1.0; P.R_ISM = 0.0; P.centre_ISM = (0.0, 0.0, 0.0)
    repartition_energie(P)
    # ---- per-grain thermal tables (init_reemission :554-640)
    cst_E = 2.0 * HP * C_LIGHT ** 2 * 4.0 * PI
    wl = lam * 1.0e-6; dwl = P.tab_delta_lambda * 1.0e-6
    B = np.zeros((n_lambda, n_T)); dB = np.zeros((n_lambda, n_T))
    for t in range(n_T):
        cw = THERMAL_CONST / float(P.tab_Temp[t]) / wl
        ok = cw < 500.0
        ce = np.exp(np.where(ok, cw, 1.0))
        B[:, t] = np.where(ok, 1.0 / ((wl ** 5) * (ce - 1.0)) * dwl, 0.0)
        dB[:, t] = np.where(ok, B[:, t] * cw * ce / (ce - 1.0), 0.0)
    Cn = P.C_abs_norm.astype(np.float64)

    def one_grain(ks):
        integ = Cn[ks] @ B                                                       # (nk, n_T)
        logE = np.where(integ > np.finfo(np.float64).tiny, np.log(np.maximum(integ, 1e-300) * cst_E), -1000.0)
        i3 = np.cumsum(Cn[ks][:, 1:, None] * dB[None, 1:, :], axis=1)            # (nk, nl-1, n_T): integ3(2:n_lambda)
        i3 = np.concatenate((np.zeros((len(ks), 1, n_T)), i3), axis=1)
        cdf_ = i3 / np.maximum(i3[:, -1:, :], 1e-300)
        return np.asfortranarray(logE), np.asfortranarray(np.transpose(cdf_, (1, 0, 2)))     # (nk, n_T), (nl, nk, n_T)
    _, P.kdB_dT_1grain_LTE_CDF = one_grain(kL)
    P.log_E_em_1grain, P.kdB_dT_1grain_nLTE_CDF = one_grain(kN)
    P.log_E_em_1grain_nRE, P.kdB_dT_1grain_nRE_CDF = one_grain(kR)
    # kabs_nLTE_CDF(grain_RE_nLTE_start-1:grain_RE_nLTE_end, n_cells, n_lambda) (dust_prop.f90:930-945)
    pkN = kN if variable else np.zeros(len(kN), int)
    cab = np.cumsum(np.einsum("kl,kc->kcl", C_abs[kN].astype(np.float64), dd[pkN, :] * ng[kN, None]), axis=0)
    cab = np.concatenate((np.zeros((1, nc, n_lambda)), cab))
    last = cab[-1:]
    P.kabs_nLTE_CDF = np.asfortranarray(np.where(last > np.finfo(np.float32).tiny, cab / np.maximum(last, 1e-300), cab))
    # ---- nRE grains: which are at quasi equilibrium in which cell, and the probabilities that follow
    rng = np.random.default_rng(seed)
    l_RE = (rng.random((n_nRE, nc)) < qre_fraction)
    P.l_RE = np.asfortranarray(l_RE.astype(np.int32))
    full = dd[(np.arange(K) if variable else np.zeros(K, int)), :] * ng[:, None]                     # (K, n_cells)
    k_tot = np.einsum("kl,kc->cl", C_abs.astype(np.float64), full)
    k_LTE = np.einsum("kl,kc->cl", C_abs[kL].astype(np.float64), full[kL])
    k_nLTE = np.einsum("kl,kc->cl", C_abs[kN].astype(np.float64), full[kN])
    k_qRE = np.einsum("kl,kc->cl", C_abs[kR].astype(np.float64), full[kR] * l_RE)
    k_RE = k_LTE + k_nLTE + k_qRE
    P.kappa_abs_RE = np.asfortranarray(k_RE * AU_TO_CM_MUM2)
    P.proba_abs_RE = np.asfortranarray(np.where(l_RE.all(0)[:, None], 1.0, k_RE / k_tot))
    P.Proba_abs_RE_LTE = np.asfortranarray(k_LTE / k_RE)
    P.Proba_abs_RE_LTE_p_nLTE = np.asfortranarray((k_LTE + k_nLTE) / k_RE)
    P.J0 = np.asfortranarray(1.0e-10 * P.volume[:, None] * np.ones((1, n_lambda)))
    P.name = "multi-grain disk (LTE + nLTE + nRE/qRE grains%s)" % (", lvariable_dust" if variable else "")
    return P
