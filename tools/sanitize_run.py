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
# round-2 neighbours of the path: dark zone, column densities, emission tables, rt1 source function and formal solution
Pd = S.ref41_like(n_photons_eq_th=10, dark_zone=False, n_rad=30, nz=16, n_rad_in=5, tau_mid=1.0e5)
Gd = api.PhotonLoop(Pd)
d = Gd.define_dark_zone(Pd.lambda_seuil, 300.0, Pd.r_grid, Pd.z_grid, [(1, Pd.n_rad)])
print("define_dark_zone", d["l_dark_zone"].sum(), d["ri_in"], d["ri_out"])
cx, cy, cz = S.cell_centres(Pd)
print("compute_column", Gd.compute_column(Pd.lambda_seuil, cx, cy, cz).max(), Gd.compute_column(1, cx, cy, cz, np.ones(Pd.n_cells)).max())
lq, cdf = Gd.init_reemission(Pd.tab_lambda, Pd.tab_delta_lambda)
print("init_reemission", lq.max(), cdf.max())
kw = dict(letape_th=0, lmono=1, lscatt_ray_tracing1=1, lsepar_pola=1, lsepar_contrib=1, RT_n_incl=2, RT_n_az=1,
          tab_u_rt=np.array([[0.0], [0.5]]), tab_v_rt=np.zeros((2, 1)), tab_w_rt=np.array([1.0, np.sqrt(0.75)]))
t = Gd.mc_photon_loop(6, 6, 10 ** 9, 20.0, **kw)
eps = Gd.init_dust_source_fct1(6, 2, 1.0e-3, np.full(Pd.n_cells, 1.0e-6), 8)
from helpers import rays_in_cells
ic, x, y, z, u, v, w = rays_in_cells(Pd, 500, seed=3)
print("rt1", np.abs(eps).sum(), Gd.integ_ray_dust(6, x, y, z, u, v, w, ic, 100.0, 8).sum())
Gd.close()
P3 = small_problems()["cyl3D"]()
G3 = api.PhotonLoop(P3)
d3 = G3.define_dark_zone(P3.lambda_seuil, 20.0, P3.r_grid, P3.z_grid, [(1, P3.n_rad)])
print("define_dark_zone 3D", d3["l_dark_zone"].sum())
G3.close()
Pg = S.multi_grain_like(n_photons_eq_th=10)
Gg = api.PhotonLoop(Pg)
print("init_reemission_grains", Gg.init_reemission_grains(Pg.tab_lambda, Pg.tab_delta_lambda, Pg.C_abs_norm, Pg.grain_nRE_start, Pg.grain_nRE_end)[0].max())
Gg.close()
print("sanitize run done")
