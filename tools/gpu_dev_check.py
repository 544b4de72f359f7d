"""Developer diagnostics on a GPU box: GPU path vs oracle, prints instead of asserting."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S, api
from oracle.binding import Oracle

def rays_in_cells(P, n, seed=0):
    rng = np.random.default_rng(seed)
    ic = rng.integers(1, P.n_cells + 1, n).astype(np.int32)
    ci, cj, ck = P.cell_map_i[ic-1], P.cell_map_j[ic-1], P.cell_map_k[ic-1]
    f = rng.uniform(0.05, 0.95, (3, n))
    if P.kind == 1:
        r = np.sqrt(P.r_lim_2[ci-1] + f[0]*(P.r_lim_2[ci]-P.r_lim_2[ci-1]))
        aj = np.abs(cj)
        z = P.z_lim[ci-1, aj-1] + f[1]*(P.z_lim[ci-1, aj]-P.z_lim[ci-1, aj-1])
        z = np.where(cj < 0, -z, z)
        if not P.l3D: z = np.where(rng.uniform(size=n) < 0.5, -z, z)
        phi = 2*np.pi*(ck-1+f[2])/P.n_az
        x, y = r*np.cos(phi), r*np.sin(phi)
    else:
        rr = (P.r_lim_3[ci-1] + f[0]*(P.r_lim_3[ci]-P.r_lim_3[ci-1]))**(1/3)
        aj = np.abs(cj)
        th = P.theta_lim[aj-1] + f[1]*(P.theta_lim[aj]-P.theta_lim[aj-1])
        th = np.where(cj < 0, -th, th)
        if not P.l3D: th = np.where(rng.uniform(size=n) < 0.5, -th, th)
        phi = 2*np.pi*(ck-1+f[2])/P.n_az
        z = rr*np.sin(th); x = rr*np.cos(th)*np.cos(phi); y = rr*np.cos(th)*np.sin(phi)
    w = rng.uniform(-1, 1, n); ph = rng.uniform(0, 2*np.pi, n)
    u = np.sqrt(1-w*w)*np.cos(ph); v = np.sqrt(1-w*w)*np.sin(ph)
    return ic, x, y, z, u, v, w

def check_geom(name, P):
    O = Oracle(P); G = api.PhotonLoop(P)
    n = 200000
    ic, x, y, z, u, v, w = rays_in_cells(P, n)
    io, ig = O.index_cell(x, y, z), G.index_cell(x, y, z)
    print(f"[{name}] index_cell: oracle==input {np.mean(io==ic):.6f}  gpu==oracle {np.mean(ig==io):.6f}")
    o = O.cross_cell(x, y, z, u, v, w, ic); g = G.cross_cell(x, y, z, u, v, w, ic)
    print(f"[{name}] cross_cell: next_cell equal {np.mean(o['next_cell']==g['next_cell']):.6f}  l bit-equal {np.mean(o['l']==g['l']):.6f}"
          f"  max rel dl {np.max(np.abs(o['l']-g['l'])/np.maximum(np.abs(o['l']),1e-300)):.3e}  x1 bit-equal {np.mean(o['x1']==g['x1']):.6f} z1 {np.mean(o['z1']==g['z1']):.6f}")
    bad = np.where(o['next_cell'] != g['next_cell'])[0][:5]
    for b in bad: print("   mismatch", ic[b], o['next_cell'][b], g['next_cell'][b], o['l'][b], g['l'][b])
    t0=time.time(); o = O.optical_length_tot(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000]); t1=time.time()
    g = G.optical_length_tot(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000]); t2=time.time()
    print(f"[{name}] optical_length_tot: steps equal {np.mean(o['n_steps']==g['n_steps']):.6f}  tau bit-equal {np.mean(o['tau_tot']==g['tau_tot']):.6f} "
          f"max rel {np.max(np.abs(o['tau_tot']-g['tau_tot'])/np.maximum(o['tau_tot'],1e-300)):.3e} lmax bit-equal {np.mean(o['lmax']==g['lmax']):.6f}  (oracle {t1-t0:.2f}s gpu {t2-t1:.2f}s)")
    # rays from outside
    rng = np.random.default_rng(5); m = 50000
    R = 3*np.sqrt(P.Rmax2)
    cz = rng.uniform(-1,1,m); ph = rng.uniform(0,2*np.pi,m)
    xs, ys, zs = R*np.sqrt(1-cz*cz)*np.cos(ph), R*np.sqrt(1-cz*cz)*np.sin(ph), R*cz
    tx, ty, tz = rng.uniform(-1,1,(3,m))*np.sqrt(P.Rmax2)*0.7
    d = np.stack([tx-xs, ty-ys, tz-zs]); d /= np.linalg.norm(d, axis=0)
    o = O.move_to_grid(xs, ys, zs, d[0], d[1], d[2]); g = G.move_to_grid(xs, ys, zs, d[0], d[1], d[2])
    print(f"[{name}] move_to_grid: intersect frac {o['lintersect'].mean():.3f} equal {np.mean(o['lintersect']==g['lintersect']):.6f} icell equal {np.mean(o['icell']==g['icell']):.6f} x bit-equal {np.mean(o['x']==g['x']):.6f}")
    tau = rng.exponential(1.0, 20000).astype(np.float32)*5
    o = O.physical_length(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000], tau)
    g = G.physical_length(P.lambda_seuil, x[:20000], y[:20000], z[:20000], u[:20000], v[:20000], w[:20000], ic[:20000], tau)
    print(f"[{name}] physical_length: sortie equal {np.mean(o['flag_sortie']==g['flag_sortie']):.6f} icell equal {np.mean(o['icell']==g['icell']):.6f} x bit-equal {np.mean(o['x']==g['x']):.6f} ltot equal {np.mean(o['ltot']==g['ltot']):.6f} u equal {np.mean(o['u']==g['u']):.6f}")
    return O, G

if __name__ == "__main__":
    t0 = time.time()
    P = S.ref41_like(n_photons_eq_th=1000, dark_zone=False)
    O = Oracle(P)
    P.l_dark_zone = S.define_dark_zone(P, P.lambda_seuil, 1500.0, O.dark_zone_walker())
    G0 = api.PhotonLoop(P)
    dz_g = S.define_dark_zone(P, P.lambda_seuil, 1500.0, G0.dark_zone_walker())
    print("dark zone: oracle", P.l_dark_zone.sum(), "gpu", dz_g.sum(), "equal", (dz_g == P.l_dark_zone).all())
    S.repartition_energie(P)
    O, G = check_geom("cyl2D", P)
    # thermal run
    for npk in (100, 1000):
        t1 = time.time(); to = O.run(n_threads=0, n_photons2=npk); t2 = time.time()
        tg = G.mc_photon_loop(1, 1, npk, 1e30, 1, False); t3 = time.time()
        ms = G.last_kernel_ms()
        print(f"thermal n2={npk}: oracle {t2-t1:.2f}s ({to.stats[0]/(t2-t1):.0f} pk/s)  gpu wall {t3-t2:.3f}s kernel {ms:.1f} ms ({tg.stats[0]/ms*1e3:.0f} pk/s)")
        print("   stats oracle", to.stats)
        print("   stats gpu   ", tg.stats)
        print("   sum sed", to.sed.sum(), tg.sed.sum(), " n_env", to.n_phot_envoyes.sum(), tg.n_phot_envoyes.sum())
        m = (to.xKJ_abs > 0) & (tg.xKJ_abs > 0)
        rel = np.abs(tg.xKJ_abs[m]-to.xKJ_abs[m])/to.xKJ_abs[m]
        print(f"   xKJ sum rel {abs(tg.xKJ_abs.sum()-to.xKJ_abs.sum())/to.xKJ_abs.sum():.4f} median cell rel {np.median(rel):.4f}  cells>0: {m.sum()}")
        print("   sed by incl (oracle)", to.sed.sum(axis=(0,2))[:5], "\n   sed by incl (gpu)   ", tg.sed.sum(axis=(0,2))[:5])
    big = 20000
    G.mc_photon_loop(1, 1, big, 1e30, 1, False); ms = G.last_kernel_ms()
    print(f"thermal n2={big}: kernel {ms:.1f} ms -> {128*big/ms*1e3:.3e} pk/s")
    for nm, PP in (("cyl3D", S.ref41_3d_like(n_photons_eq_th=100, n_rad=30, nz=10, n_az=12, n_rad_in=4, tau_mid=100.)),
                   ("sph2D", S.spherical_shell(n_photons_eq_th=100)),
                   ("sph3D", S.spherical_shell(n_photons_eq_th=100, n_az=8, l3D=True))):
        O2, G2 = check_geom(nm, PP)
        to = O2.run(n_threads=0, n_photons2=100); tg = G2.mc_photon_loop(1, 1, 100, 1e30, 1, False)
        print(f"[{nm}] thermal: stats oracle {to.stats}\n            stats gpu    {tg.stats}\n    xKJ sum rel {abs(tg.xKJ_abs.sum()-to.xKJ_abs.sum())/to.xKJ_abs.sum():.4f}  sed {to.sed.sum()} {tg.sed.sum()}")
    print("total", time.time()-t0)
