"""GPU against the oracle run with 1 and with all host threads, G4 / G5 at full size (usage: fullsize_diag.py g4|g5 n2).
The reference's running temperature estimate uses the thread's OWN tally x nb_proc (thermal_emission.f90:668 with id), so at
few packets per cell its answer depends on the thread count; the device reads the global running tally (= nb_proc 1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
from oracle.binding import Oracle

which, n2 = sys.argv[1], int(sys.argv[2])
if which == "g4":
    P = S.ref41_3d_like(n_photons_eq_th=n2, tau_mid=1.0e3, n_rad=100, nz=50, n_az=72)
else:
    P = S.voronoi_sph_disk(n_points=1000000, n_photons_eq_th=n2, tau_mid=1.0e3, cache=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data_cache", "g5_1000000.npz"))
G = api.PhotonLoop(P)
tg = G.mc_photon_loop(1, 1, n2, 1.0e30, 1, False)
print("gpu ms", G.last_kernel_ms(), tg.stats[:7])
G.close()
O = Oracle(P, fast=True)


def cmp(a, b, tag):
    no, ng = a.n_phot_sed[:, :, 0], b.n_phot_sed[:, :, 0]
    m = (no + ng) > 50
    z = (ng[m] - no[m]) / np.sqrt(no[m] + ng[m])
    print(tag, "sed bins", m.sum(), "frac<3 %.4f" % np.mean(np.abs(z) < 3), "mean z %.3f" % z.mean(), "max |z| %.2f" % np.abs(z).max(),
          "| absorbed %.4f steps %.4f inter %.4f" % (b.xKJ_abs.sum() / a.xKJ_abs.sum() - 1, b.stats[1] / a.stats[1] - 1, b.stats[2] / a.stats[2] - 1))
    sa, sb = no.sum(axis=1), ng.sum(axis=1)
    print("   per-lambda ratio", np.round(sb / np.maximum(sa, 1), 3)[14:])


res = {}
for nt in (0, 1):
    t = time.time(); res[nt] = O.run(n_threads=nt, n_photons2=n2); print("oracle threads", nt, "%.1f s" % (time.time() - t))
cmp(res[1], tg, "gpu vs oracle(1 thread)   ")
cmp(res[0], tg, "gpu vs oracle(all threads)")
cmp(res[1], res[0], "oracle(all) vs oracle(1)  ")
