"""Where should a thermal call switch from the packet-per-warp kernel alone to the packet-per-lane kernel + stragglers?
The switch is at 4 x 128 x blocks / max_inflight_fraction packets: the same budgets are run with the default fraction
(packet-per-lane path above 1.2e6 packets) and with a small fraction (packet-per-warp kernel alone)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = G.define_dark_zone(P.lambda_seuil, 1500.0, P.r_grid, P.z_grid, [(1, P.n_rad)])["l_dark_zone"]
S.repartition_energie(P); G.upload_emission(P)
flags = dict(lsepar_pola=1, lsepar_contrib=1)
G.mc_photon_loop(1, 1, 200, **flags)
for n2 in (10000, 20000, 30000, 50000):
    P.n_photons_eq_th = n2; S.repartition_energie(P); G.upload_emission(P)
    for frac, tag in ((0.0, "default (lane path)"), (1.0 / 400.0, "packet-per-warp kernel alone")):
        best = 1e30
        for rep in range(2):
            t0 = time.perf_counter()
            t = G.mc_photon_loop(1, 1, n2, call_index=rep, max_inflight_fraction=frac, **flags)
            best = min(best, time.perf_counter() - t0)
        d = G.debug_counters()
        print("%8d packets  %-30s %.1f ms  launches %d parked %d" % (128 * n2, tag, 1e3 * best, d["launches"], d["parked"]))
