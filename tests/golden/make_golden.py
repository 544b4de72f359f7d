"""Generates tests/golden/oracle_golden.npz: the oracle's own regression pins
(the reference ships no golden vectors for these routines).
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import rays_in_cells, small_problems  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

out = {}
for name in ("cyl2D", "cyl3D", "sph2D", "sph3D"):
    P = small_problems()[name]()
    O = Oracle(P)
    ic, x, y, z, u, v, w = rays_in_cells(P, 512, seed=11)
    c = O.cross_cell(x, y, z, u, v, w, ic)
    r = O.optical_length_tot(P.lambda_seuil, x, y, z, u, v, w, ic)
    out[f"{name}_next_cell"] = c["next_cell"]
    out[f"{name}_l"] = c["l"]
    out[f"{name}_n_steps"] = r["n_steps"]
    out[f"{name}_tau"] = r["tau_tot"]
P = small_problems()["cyl2D"]()
t = Oracle(P).run(n_threads=1, n_photons2=5)
out["thermal_stats"] = t.stats
out["thermal_xKJ"] = t.xKJ_abs
out["thermal_sed"] = t.sed
# per-grain branches and the complete capteur (single thread: deterministic)
from mcfost_b200 import synthetic as S  # noqa: E402
PG = S.multi_grain_like(n_photons_eq_th=5, n_rad=10, nz=6, n_rad_in=2, tau_mid=10.0)
OG = Oracle(PG)
t = OG.run(n_threads=1, xJ=True, n_photons2=5, lonly_LTE=0, lRE_nLTE=1, lnRE=1, lxJ_abs_step1=1)
out["mixed_stats"] = t.stats
out["mixed_xKJ"] = t.xKJ_abs
out["mixed_xT_1grain"] = t.xT_ech_1grain
out["mixed_xT_1grain_nRE"] = t.xT_ech_1grain_nRE
out["mixed_E_abs_nRE"] = t.E_abs_nRE
t = OG.run(n_threads=1, letape_th=0, lmono=1, lambda_in=6, p_lambda_in=6, n_photons2=10 ** 9, n_phot_lim=5.0,
           lscattering_method1=1, lsepar_pola=1)
out["method1_stats"] = t.stats
out["method1_sed"] = t.sed
out["method1_sed_q"] = t.sed_q
t = OG.run(n_threads=1, letape_th=0, lmono=1, lmono0=1, loutput_mc=1, lambda_in=6, p_lambda_in=6, n_photons2=5,
           npix_x=8, npix_y=8, map_size=300.0, N_thet=3, N_phi=1, lsepar_pola=1, lsepar_contrib=1, l_sym_ima=1)
out["maps_stokes"] = t.stokes_map
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
