"""Speed of a thermal call against the in-flight window (max_inflight_fraction): G1 at 2.56e6 and 1.28e7 packets, G5 at 2.56e6."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
which = sys.argv[1]
if which == "g1":
    P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
    budgets = (20000, 100000); flags = dict(lsepar_pola=1, lsepar_contrib=1)
else:
    P = S.voronoi_sph_disk(n_points=1000000, n_photons_eq_th=20000, tau_mid=1.0e3,
                           cache=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data_cache", "g5_1000000.npz"))
    budgets = (20000,); flags = dict(lsepar_pola=0, lsepar_contrib=0)
G = api.PhotonLoop(P)
if which == "g1":
    P.l_dark_zone = G.define_dark_zone(P.lambda_seuil, 1500.0, P.r_grid, P.z_grid, [(1, P.n_rad)])["l_dark_zone"]
G.mc_photon_loop(1, 1, 200, **flags)
for n2 in budgets:
    P.n_photons_eq_th = n2; S.repartition_energie(P); G.upload_emission(P)
    for frac in (1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0 / 2):
        best = 1e30
        for rep in range(2):
            t0 = time.perf_counter(); t = G.mc_photon_loop(1, 1, n2, call_index=rep, max_inflight_fraction=frac, **flags); best = min(best, time.perf_counter() - t0)
        d = G.debug_counters()
        print("%s %9d packets  fraction 1/%-3d %.1f ms  dry %.1f  lane end %.1f  launches %d  fill FLY %.1f" % (which, 128 * n2, round(1 / frac), 1e3 * best, d["steady_ms"], d["main_end_ms"], d["launches"], d["chunk_fill"]["FLY"]))
