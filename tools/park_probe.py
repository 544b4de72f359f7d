"""Timeline of one thermal call with the straggler hand-over (usage: park_probe.py n2 overlap_sms [pola])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ov = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pola = int(sys.argv[3]) if len(sys.argv) > 3 else 0
P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
G.mc_photon_loop(1, 1, 200)
G.set_overlap(ov)
for rep in range(int(os.environ.get("REPS", "2"))):
    t = G.mc_photon_loop(1, 1, n2, call_index=rep, lsepar_pola=pola, lsepar_contrib=pola)
    d = G.debug_counters()
    print(f"n2={n2} ov={ov} park_live={os.environ.get('MCB_PARK_LIVE','32')} event-ms {G.last_kernel_ms():.0f}  dry {d['steady_ms']:.0f}  main_end {d['main_end_ms']:.0f}  straggler_end {d['straggler_end_ms']:.0f}  parked {d['parked']:.0f}  packets {t.stats[0]:.0f}", flush=True)
