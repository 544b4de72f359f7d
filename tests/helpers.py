"""Shared test helpers: seeded rays inside cells of the synthetic grids."""
import numpy as np


def rays_in_cells(P, n, seed=0):
    """Random points strictly inside random real cells + isotropic directions."""
    rng = np.random.default_rng(seed)
    ic = rng.integers(1, P.n_cells + 1, n).astype(np.int32)
    ci, cj, ck = P.cell_map_i[ic - 1], P.cell_map_j[ic - 1], P.cell_map_k[ic - 1]
    f = rng.uniform(0.05, 0.95, (3, n))
    aj = np.abs(cj)
    phi = 2 * np.pi * (ck - 1 + f[2]) / P.n_az
    if P.kind == 1:
        r = np.sqrt(P.r_lim_2[ci - 1] + f[0] * (P.r_lim_2[ci] - P.r_lim_2[ci - 1]))
        z = P.z_lim[ci - 1, aj - 1] + f[1] * (P.z_lim[ci - 1, aj] - P.z_lim[ci - 1, aj - 1])
        z = np.where(cj < 0, -z, z)
        if not P.l3D:
            z = np.where(rng.uniform(size=n) < 0.5, -z, z)
        x, y = r * np.cos(phi), r * np.sin(phi)
    else:
        rr = (P.r_lim_3[ci - 1] + f[0] * (P.r_lim_3[ci] - P.r_lim_3[ci - 1])) ** (1 / 3)
        th = P.theta_lim[aj - 1] + f[1] * (P.theta_lim[aj] - P.theta_lim[aj - 1])
        th = np.where(cj < 0, -th, th)
        if not P.l3D:
            th = np.where(rng.uniform(size=n) < 0.5, -th, th)
        z = rr * np.sin(th)
        x, y = rr * np.cos(th) * np.cos(phi), rr * np.cos(th) * np.sin(phi)
    w = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    u, v = np.sqrt(1 - w * w) * np.cos(ph), np.sqrt(1 - w * w) * np.sin(ph)
    return ic, x, y, z, u, v, w


def rays_from_outside(P, m, seed=5):
    rng = np.random.default_rng(seed)
    R = 3 * np.sqrt(P.Rmax2)
    cz = rng.uniform(-1, 1, m)
    ph = rng.uniform(0, 2 * np.pi, m)
    xs, ys, zs = R * np.sqrt(1 - cz * cz) * np.cos(ph), R * np.sqrt(1 - cz * cz) * np.sin(ph), R * cz
    t = rng.uniform(-1, 1, (3, m)) * np.sqrt(P.Rmax2) * 0.7
    d = np.stack([t[0] - xs, t[1] - ys, t[2] - zs])
    d /= np.linalg.norm(d, axis=0)
    return xs, ys, zs, d[0], d[1], d[2]


def small_problems():
    """The four structured-grid flavours at sizes the oracle handles in seconds."""
    from mcfost_b200 import synthetic as S
    return {
        "cyl2D": lambda: S.ref41_like(n_photons_eq_th=100, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3),
        "cyl3D": lambda: S.ref41_3d_like(n_photons_eq_th=100, n_rad=30, nz=10, n_az=12, n_rad_in=4, tau_mid=100.0),
        "sph2D": lambda: S.spherical_shell(n_photons_eq_th=100),
        "sph3D": lambda: S.spherical_shell(n_photons_eq_th=100, n_az=8, l3D=True),
    }
