"""One thermal photon-loop launch on G1 for ncu (usage: prof_run.py n2 [tau_mid] [pola] [mrw])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
tau = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0e5
pola = int(sys.argv[3]) if len(sys.argv) > 3 else 0
mrw = int(sys.argv[4]) if len(sys.argv) > 4 else 0
P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False, tau_mid=tau)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
flags = dict(lsepar_pola=pola, lsepar_contrib=pola)
if mrw:
    flags["lMRW"] = 1
t = G.mc_photon_loop(1, 1, n2, **flags)
print("kernel ms", G.last_kernel_ms(), "stats", t.stats)
d = G.debug_counters()
print("steady ms", d["steady_ms"], "fill", d["chunk_fill"])
