"""The SED step of G1 at the .para budget: one mc_photon_loop call per wavelength (run_sed_mc, dust_transfer.f90:826-1045:
lmono, forced scattering, rt1 accumulators, chunks end when n_photons2 = 10 packets were RECEIVED in detector bin capt_sup or
nbre_photons_lambda x ... sent), GPU (through the C ABI, host buffers) against the oracle on the host cores.
usage: sed_step.py [n_lambda_max]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
from oracle.binding import Oracle

nl_max = int(sys.argv[1]) if len(sys.argv) > 1 else 50
P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
G = api.PhotonLoop(P)
P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G.dark_zone_walker())
S.repartition_energie(P); G.upload_dark_zone(P.l_dark_zone); G.upload_emission(P)
O = Oracle(P, fast=True)
incl = np.deg2rad([45.0, 60.0, 75.0])
kw = dict(letape_th=0, lmono=1, lscatt_ray_tracing1=1, lsepar_pola=1, lsepar_contrib=1, RT_n_incl=3, RT_n_az=1,
          tab_u_rt=np.sin(incl).reshape(3, 1), tab_v_rt=np.zeros((3, 1)), tab_w_rt=np.cos(incl))
n_photons2, n_phot_lim = 10, 1.28e3 * 1000.0 / 128.0       # read_param.f90:145-149: nbre_photons_lambda = 1.28e3 -> per-chunk limits
lams = list(range(1, min(P.n_lambda, nl_max) + 1))
G.mc_photon_loop(lams[0], lams[0], n_photons2, n_phot_lim, 1, False, **kw)      # warm-up
tg = {}
t0 = time.perf_counter()
for i, l in enumerate(lams):
    tg[l] = G.mc_photon_loop(l, l, n_photons2, n_phot_lim, 1, False, call_index=i, reset_tallies=1, **kw)
t_gpu = time.perf_counter() - t0
ntf = 8
n_xI = 45 * 2 * ntf * 3 * P.n_cells
nthr = len(os.sched_getaffinity(0))
O.run(n_threads=nthr, n_xI=n_xI, lambda_in=lams[0], p_lambda_in=lams[0], n_photons2=n_photons2, n_phot_lim=n_phot_lim, **kw)
to = {}
t0 = time.perf_counter()
for i, l in enumerate(lams):
    to[l] = O.run(n_threads=nthr, n_xI=n_xI, lambda_in=l, p_lambda_in=l, n_photons2=n_photons2, n_phot_lim=n_phot_lim, call_index=i, **kw)
t_cpu = time.perf_counter() - t0
pk_g = sum(t.stats[0] for t in tg.values()); pk_o = sum(t.stats[0] for t in to.values())
print("SED step, %d wavelengths: GPU %.3f s (%.0f packets, %.3e pk/s)   oracle on %d threads %.3f s (%.0f packets, %.3e pk/s)" % (len(lams), t_gpu, pk_g, pk_g / t_gpu, nthr, t_cpu, pk_o, pk_o / t_cpu))
for l in lams[::7]:
    a, b = tg[l], to[l]
    sa, sb = a.sed[l - 1].sum() / max(a.n_phot_envoyes[l - 1], 1), b.sed[l - 1].sum() / max(b.n_phot_envoyes[l - 1], 1)
    print("  lambda %2d: packets sent GPU %8.0f oracle %8.0f   escaping fraction x energy GPU %.4e oracle %.4e" % (l, a.stats[0], b.stats[0], sa, sb))
