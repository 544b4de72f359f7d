"""The SED step of G1 (see sed_step.py) with the wavelengths dealt over several handles: each handle's call is launched
asynchronously (mcfost_b200_launch), the calls of a group run concurrently on the GPU (every call sizes its grid from its
budget), then each is waited for and downloaded.  What a caller that pipelines the wavelength loop of run_sed_mc gets.
usage: sed_step_handles.py [n_handles ...]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api

counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
G0 = api.PhotonLoop(P)
P.l_dark_zone = G0.define_dark_zone(P.lambda_seuil, 1500.0, P.r_grid, P.z_grid, [(1, P.n_rad)])["l_dark_zone"]
S.repartition_energie(P)
G0.close()
incl = np.deg2rad([45.0, 60.0, 75.0])
kw = dict(letape_th=0, lmono=1, lscatt_ray_tracing1=1, lsepar_pola=1, lsepar_contrib=1, RT_n_incl=3, RT_n_az=1,
          tab_u_rt=np.sin(incl).reshape(3, 1), tab_v_rt=np.zeros((3, 1)), tab_w_rt=np.cos(incl))
n_photons2, n_phot_lim = 10, 1.28e3 * 1000.0 / 128.0
lams = list(range(1, P.n_lambda + 1))
ref = None
for nh in counts:
    H = [api.PhotonLoop(P) for _ in range(nh)]
    for g in H:
        g.mc_photon_loop(1, 1, n_photons2, n_phot_lim, 1, False, **kw)      # warm-up
    sed = np.zeros(P.n_lambda); sent = np.zeros(P.n_lambda)
    t0 = time.perf_counter()
    for base in range(0, len(lams), nh):
        group = lams[base:base + nh]
        runs = [H[k].launch(l, l, n_photons2, n_phot_lim, 1, call_index=l, reset_tallies=1, **kw) for k, l in enumerate(group)]
        for k, l in enumerate(group):
            H[k].sync()
            t = H[k].download(runs[k])
            sed[l - 1] = t.sed[l - 1].sum(); sent[l - 1] = t.n_phot_envoyes[l - 1]
    dt = time.perf_counter() - t0
    for g in H:
        g.close()
    esc = sed / np.maximum(sent, 1)
    if ref is None:
        ref = esc
    print("%d handle(s): 50 wavelengths in %.3f s, %.3e packets, escaping energy per packet vs 1 handle: max rel diff %.3f" %
          (nh, dt, sent.sum(), np.abs(esc / ref - 1).max()))
