"""Small calls through every kernel variant, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api


P = S.multi_grain_like(n_photons_eq_th=20, n_rad=12, nz=8, n_rad_in=3, tau_mid=20.0)
G = api.PhotonLoop(P)
t = G.mc_photon_loop(1, 1, 20)                                                     # thermal variant
print("thermal", t.stats[:7])
t = G.mc_photon_loop(1, 1, 2000, lsepar_pola=1, max_inflight_fraction=1.0)           # packet-per-lane kernel, parking, adopt launch
print("three launches", t.stats[:7], G.debug_counters()["parked"], G.debug_counters()["launches"])
Pm = S.ref41_like(n_photons_eq_th=20, dark_zone=False, n_rad=12, nz=8, n_rad_in=3, tau_mid=3.0e5)
Gm = api.PhotonLoop(Pm)
t = Gm.mc_photon_loop(1, 1, 20, lMRW=1, gamma_MRW=2.0)                               # modified random walk in the packet-per-warp kernel
print("mrw", t.stats[:10])
Gm.close()
print("closest wall", G.distance_to_closest_wall(np.arange(1, 9), np.full(8, 3.0), np.zeros(8), np.full(8, 0.1)))
G.set_overlap(2, 2)
t = G.mc_photon_loop(1, 1, 2000, lsepar_pola=1, max_inflight_fraction=1.0)           # reserved SMs, high-priority adopt launch
print("overlap", t.stats[:7], G.debug_counters()["parked"])
G.set_overlap(0)
t = G.mc_photon_loop(6, 6, 10 ** 9, 20.0, letape_th=0, lmono=1, lsepar_pola=1, lscatt_ray_tracing1=1, RT_n_incl=2, RT_n_az=1,
                     tab_u_rt=np.array([[0.0], [0.5]]), tab_v_rt=np.zeros((2, 1)), tab_w_rt=np.array([1.0, np.sqrt(0.75)]))   # generic variant, rt1
print("sed rt1", t.stats[:7])
t = G.mc_photon_loop(6, 6, 30, letape_th=0, lmono=1, lmono0=1, lscatt_ray_tracing2=1, lsepar_pola=1)                         # rt2
print("image rt2", t.stats[:7])
t = G.mc_photon_loop(6, 6, 30, letape_th=0, lmono=1, lmono0=1, loutput_mc=1, npix_x=8, npix_y=8, map_size=300.0, lorigine=1,
                     l_sym_ima=1, lsepar_pola=1, lsepar_contrib=1)                                                            # extras: maps
print("maps", t.stats[:7], t.stokes_map.sum())
t = G.mc_photon_loop(1, 1, 20, lonly_LTE=0, lRE_nLTE=1, lnRE=1, lxJ_abs_step1=1, lscattering_method1=1, lxN_abs=1,
                     low_mem_th_emission=1, low_mem_th_emission_nLTE=1)                                                       # extras: grains
print("grains", t.stats[:7], t.E_abs_nRE)
print("T", G.temp_finale().max(), G.temp_finale_nlte().max())
G.close()
Pi = S.ref41_like(n_photons_eq_th=10, dark_zone=False, n_rad=12, nz=8, n_rad_in=3, tau_mid=1.0e2)
Pi.R_ISM = 1.2 * np.sqrt(Pi.Rmax2 + Pi.zmaxmax ** 2); Pi.centre_ISM = (0.0, 0.0, 0.0)
GI = api.PhotonLoop(Pi)
t = GI.mc_photon_loop(10, 10, 30, 1.0e30, 1, False, letape_th=0, lmono=1, lISM_loop=1, lxJ_abs=1)                             # ISM side loop
print("ism loop", t.stats[:7])
GI.close()
M = api.MultiPhotonLoop(Pi, 1)
t = M.mc_photon_loop(1, 1, 20)
print("multi handle", t.stats[:7], M.temp_finale().max())
M.close()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import small_problems
for name in ("sph2D", "cyl3D", "sph3D"):
    Pg = small_problems()[name]()
    Gg = api.PhotonLoop(Pg)
    print(name, Gg.mc_photon_loop(1, 1, 10).stats[:7], Gg.mc_photon_loop(1, 1, 400, max_inflight_fraction=1.0).stats[:7])
    Gg.close()
V = S.voronoi_disk(n_points=300, n_photons_eq_th=10)
GV = api.PhotonLoop(V)
print("voronoi", GV.mc_photon_loop(1, 1, 10).stats[:7])
GV.close()
print("sanitize run done")
