"""Seeded rays + the grid tables of the cyl2D test problem, in the stream layout ref_harness.f90 reads."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rays_in_cells, small_problems


def write(path, name="cyl2D", n=20000, seed=101):
    P = small_problems()[name]()
    ic, x, y, z, u, v, w = rays_in_cells(P, n, seed=seed)
    with open(path, "wb") as f:
        np.array([P.n_rad, P.nz, P.n_az, int(P.l3D), n], np.int32).tofile(f)
        for a in (P.r_lim, P.r_lim_2, P.r_lim_3):
            np.asarray(a, np.float64).tofile(f)
        np.asfortranarray(np.asarray(P.z_lim, np.float64)).ravel(order="F").tofile(f)
        np.asarray(P.zmax, np.float64).tofile(f)
        for a in (P.tan_phi_lim, P.sin_phi_lim, P.cos_phi_lim):
            np.asarray(a, np.float64).tofile(f)
        np.array([P.Rmax2, P.zmaxmax], np.float64).tofile(f)
        for a in (x, y, z, u, v, w):
            np.asarray(a, np.float64).tofile(f)
        np.asarray(ic, np.int32).tofile(f)
    return P, (ic, x, y, z, u, v, w)


if __name__ == "__main__":
    write(sys.argv[1])
