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
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz"), **out)
print("written", {k: v.shape for k, v in out.items()})
