"""Per-event latency of a single packet: launches of ONE packet (n_photons_loop = 1, n_photons2 = 1) on G1.
The kernel time of such a launch is (events of that packet) x (latency per event) + a fixed start-up cost,
so a sweep over call_index (= different packets) separates the two.  Also: 128 packets (one per chunk).
usage: latency_probe.py [n_calls] [pola] [mrw]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pola = int(sys.argv[2]) if len(sys.argv) > 2 else 1
mrw = int(sys.argv[3]) if len(sys.argv) > 3 else 0
P = S.ref41_like(n_photons_eq_th=1000000, dark_zone=False)       # L_packet_th of the 1.28e8-packet budget
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
flags = dict(lsepar_pola=pola, lsepar_contrib=pola)
if mrw:
    flags["lMRW"] = 1
# a warm tally first (the trapped packets' temperature iteration is cheaper on a warm tally)
G.mc_photon_loop(1, 1, 20000, n_photons_loop=128, **flags)
rows = []
prev = G.download().stats.copy()      # stats accumulate with reset_tallies = 0: per-call values are differences
for c in range(n_calls):
    t = G.mc_photon_loop(1, 1, 1, n_photons_loop=1, call_index=1000 + c, reset_tallies=0, **flags)
    ms = G.last_kernel_ms()
    d = t.stats - prev; prev = t.stats.copy()
    rows.append((d[1], d[2], d[3], d[4], ms))
    t = None
r = np.array(rows)
ev = r[:, 0] + r[:, 1]
order = np.argsort(ev)
print("single-packet launches: steps interactions scatt abs ms  (sorted by events)")
for k in list(order[:3]) + list(order[-12:]):
    print("  %8d %8d %8d %8d  %9.3f ms   %.3f us/event  %.3f us/interaction" % (r[k, 0], r[k, 1], r[k, 2], r[k, 3], r[k, 4], 1e3 * r[k, 4] / max(ev[k], 1), 1e3 * r[k, 4] / max(r[k, 1], 1)))
big = ev > 2000
if big.sum() >= 3:
    A = np.stack([r[big, 0] - r[big, 1], r[big, 2], r[big, 3], np.ones(big.sum())], 1)     # free steps, scatterings, absorptions, const
    coef, *_ = np.linalg.lstsq(A, r[big, 4] * 1e3, rcond=None)
    print("least squares over %d packets with > 2000 events: %.3f us per extra cell step, %.3f us per scattering (+ its step), %.3f us per absorption (+ its step), %.1f us fixed" % (big.sum(), *coef))
print("fixed cost of an (almost) empty launch: median %.3f ms" % np.median(r[ev < 50, 4]) if (ev < 50).any() else "")
# 128 packets at once
for rep in range(3):
    t = G.mc_photon_loop(1, 1, 1, n_photons_loop=128, call_index=5000 + rep, reset_tallies=0, **flags)
    d = t.stats - prev; prev = t.stats.copy()
    print("128 packets: events %d  kernel %.3f ms" % (d[1] + d[2], G.last_kernel_ms()))
for n2 in (10, 100, 1000, 10000):
    t = G.mc_photon_loop(1, 1, n2, n_photons_loop=128, call_index=6000 + n2, reset_tallies=0, **flags)
    d = t.stats - prev; prev = t.stats.copy()
    print("%d packets: events %d  kernel %.3f ms  %.3e packets/s" % (128 * n2, d[1] + d[2], G.last_kernel_ms(), 128 * n2 / G.last_kernel_ms() * 1e3))
