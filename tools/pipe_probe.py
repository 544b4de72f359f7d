"""Device timeline of alternating calls on two handles (usage: pipe_probe.py n2 overlap_sms steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S, api
n2 = int(sys.argv[1]); ov = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
NH = int(sys.argv[4]) if len(sys.argv) > 4 else 2
P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False)
G = [api.PhotonLoop(P) for _ in range(NH)]
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G[0].dark_zone_walker())
S.repartition_energie(P)
for g in G:
    g.upload_dark_zone(P.l_dark_zone); g.upload_emission(P); g.mc_photon_loop(1, 1, 200); g.set_overlap(ov, max(1, ov // max(1, NH - 1)) if ov else 0)
rows = []
base = None
for i in range(steps + NH):
    g = G[i % NH]
    if i >= NH:
        d = g.debug_counters()          # previous call of this handle (syncs its stream, as the next launch would anyway)
        rows.append((i - NH, d))
    if i < steps:
        g.launch(1, 1, n2, 1.0e30, 1, call_index=i, reset_tallies=1, lsepar_pola=1, lsepar_contrib=1)
t00 = rows[0][1]["t0_ms"]
for i, d in rows:
    t0 = d["t0_ms"] - t00
    print(f"step {i} handle {i%NH}: start {t0:8.0f}  dry {t0+d['steady_ms']:8.0f}  main_end {t0+d['main_end_ms']:8.0f}  straggler_start {t0+d['straggler_start_ms'] if d['straggler_start_ms']>0 else -1:8.0f}  straggler_end {t0+d['straggler_end_ms'] if d['straggler_end_ms']>0 else -1:8.0f}  parked {d['parked']:.0f}")
