"""One thermal photon-loop call on the G5 Voronoi mesh for ncu (usage: prof_run_g5.py n2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P = S.voronoi_sph_disk(n_points=1000000, n_photons_eq_th=n2, tau_mid=1.0e3,
                       cache=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data_cache", "g5_1000000.npz"))
G = api.PhotonLoop(P)
t = G.mc_photon_loop(1, 1, n2, lsepar_pola=0, lsepar_contrib=0)
print("kernel ms", G.last_kernel_ms(), "stats", t.stats[:8])
d = G.debug_counters()
print("steady ms", d["steady_ms"], "fill", d["chunk_fill"])
