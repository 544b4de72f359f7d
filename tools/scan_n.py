"""Kernel time vs packet count: separates the throughput phase from the latency-bound tail."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
ns = [int(a) for a in sys.argv[1:]] or [1000, 20000, 100000]
P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
G.mc_photon_loop(1, 1, 200)
for n2 in ns:
    P.n_photons_eq_th = n2; S.repartition_energie(P); G.upload_emission(P)      # L_packet_th = L_tot / n_packets
    for rep in range(int(os.environ.get("REPS", "2"))):
        t = G.mc_photon_loop(1, 1, n2, call_index=rep)
        ms = G.last_kernel_ms()
        print(f"n2={n2} packets={128*n2} kernel {ms:.1f} ms  {128*n2/ms*1e3:.3e} pk/s  steps/s {t.stats[1]/ms*1e3:.3e}  int/s {t.stats[2]/ms*1e3:.3e}  steps/pk {t.stats[1]/t.stats[0]:.1f} int/pk {t.stats[2]/t.stats[0]:.1f}", flush=True)
        d = G.debug_counters()
        print(f"    steady {d['steady_ms']:.1f} ms ({128*n2/max(d['steady_ms'],1e-9)*1e3:.3e} pk/s in steady state)  drain {d['kernel_ms']-d['steady_ms']:.1f} ms  fill " + " ".join(f"{k}={v:.1f}" for k, v in d['chunk_fill'].items()), flush=True)
