"""Per-cell crossing counts of the G1 thermal step (xN_abs, radiation_field.f90:53) -> gpurun_out/g1_hits.f64, the
weights of tools/l2_atomic_peak.cu's second distribution."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
t = G.mc_photon_loop(1, 1, n2, lxN_abs=1)
x = np.ascontiguousarray(t.xN_abs.ravel(), np.float64)
os.makedirs("gpurun_out", exist_ok=True)
x.tofile("gpurun_out/g1_hits.f64")
top = np.sort(x)[::-1]
print("cells", len(x), "crossings", x.sum(), "steps", t.stats[1], "share of the 10 / 100 / 1000 most crossed cells: %.3f %.3f %.3f" % (top[:10].sum() / x.sum(), top[:100].sum() / x.sum(), top[:1000].sum() / x.sum()))
