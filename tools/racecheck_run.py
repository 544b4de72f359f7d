"""define_dark_zone only (one block with shared-memory hand-offs between the columns), for compute-sanitizer --tool racecheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
P = S.ref41_like(n_photons_eq_th=10, dark_zone=False, n_rad=30, nz=16, n_rad_in=5, tau_mid=1.0e5)
G = api.PhotonLoop(P)
d = G.define_dark_zone(P.lambda_seuil, 300.0, P.r_grid, P.z_grid, [(1, P.n_rad)])
print("dark cells", d["l_dark_zone"].sum(), d["ri_in"], d["ri_out"], d["l_is_dark_zone"])
G.close()
