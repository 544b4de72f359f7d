"""Kernel time vs packet count: separates the throughput phase from the latency-bound tail.
usage: scan_n.py [pola] [mrw] n2 n2 ..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
pola = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mrw = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ns = [int(a) for a in sys.argv[3:]] or [1000, 10000, 100000, 1000000]
P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
flags = dict(lsepar_pola=pola, lsepar_contrib=pola, lMRW=mrw)
G.mc_photon_loop(1, 1, 200, **flags)
for n2 in ns:
    P.n_photons_eq_th = n2; S.repartition_energie(P); G.upload_emission(P)      # L_packet_th = L_tot / n_packets
    for rep in range(int(os.environ.get("REPS", "2"))):
        t0 = time.perf_counter()
        t = G.mc_photon_loop(1, 1, n2, call_index=rep, **flags)
        wall = (time.perf_counter() - t0) * 1e3
        ms = G.last_kernel_ms()
        d = G.debug_counters()
        print(f"n2={n2} packets={128*n2} device {ms:.1f} ms  wall {wall:.1f} ms  {128*n2/ms*1e3:.3e} pk/s  steps/pk {t.stats[1]/t.stats[0]:.1f} int/pk {t.stats[2]/t.stats[0]:.1f} mrw steps/pk {t.stats[9]/t.stats[0]:.2f}"
              f" | dry {d['steady_ms']:.1f}  main_end {d['main_end_ms']:.1f}  tail_start {d['straggler_start_ms']:.1f} tail_end {d['straggler_end_ms']:.1f} parked {d['parked']:.0f}", flush=True)
