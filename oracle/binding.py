"""ctypes binding of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Nothing under mcfost_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from mcfost_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(fast_native=False):
    """Compile the oracle (g++).  fast_native=True rebuilds the timing flavour
    with -march=native on the machine that will run it."""
    args = ["make", "-s", "-C", HERE]
    if fast_native:
        subprocess.run(["rm", "-f", os.path.join(BUILD, "liboracle_fast.so")], check=False)
        args.append("MARCH=-march=native")
    subprocess.run(args, check=True)


def _load(name):
    path = os.path.join(BUILD, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_last_error.argtypes = [C.c_void_p]
    for fn in ("oracle_destroy", "oracle_set_grid", "oracle_set_dark_zone", "oracle_set_opacity",
               "oracle_set_emission", "oracle_set_grains", "oracle_n_cells_tot"):
        getattr(lib, fn).argtypes = [C.c_void_p] + ([C.c_void_p] if fn not in ("oracle_destroy", "oracle_n_cells_tot") else [])
    lib.oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int64]
    lib.oracle_distance_to_closest_wall.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 5
    lib.oracle_mrw_tables.argtypes = [C.c_void_p] * 4
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _map_args(P, r):
    kw = abi.map_tally_args(r)
    if "n_xN" in kw and kw["n_xN"] is None:
        kw["n_xN"] = P.n_lambda
    return kw


class Oracle:
    """Reference-order CPU implementation of mc_photon_loop and its callees."""

    def __init__(self, P, fast=False):
        self.lib = _load("liboracle_fast.so" if fast else "liboracle.so")
        self.h = C.c_void_p(self.lib.oracle_create())
        self.P = P
        self._g = abi.make_grid(P)
        self._o = abi.make_opacity(P)
        self._e = abi.make_emission(P) if hasattr(P, "prob_E_cell") else None
        self._check(self.lib.oracle_set_grid(self.h, self._g.ref()))
        self._check(self.lib.oracle_set_opacity(self.h, self._o.ref()))
        if self._e is not None:
            self._check(self.lib.oracle_set_emission(self.h, self._e.ref()))
        self._gr = None
        if hasattr(P, "n_grains_tot"):
            self._gr = abi.make_grains(P)
            self._check(self.lib.oracle_set_grains(self.h, self._gr.ref()))
        self.set_dark_zone(getattr(P, "l_dark_zone", None))

    def __del__(self):
        try:
            self.lib.oracle_destroy(self.h)
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle error {rc}: {self.lib.oracle_last_error(self.h).decode()}")

    def set_dark_zone(self, dz):
        self._dz = None if dz is None else np.ascontiguousarray(dz, np.int32)
        self._check(self.lib.oracle_set_dark_zone(self.h, _p(self._dz)))

    def set_emission(self, P):
        self._e = abi.make_emission(P)
        self._check(self.lib.oracle_set_emission(self.h, self._e.ref()))

    @property
    def n_cells_tot(self):
        return self.lib.oracle_n_cells_tot(self.h)

    def cell_maps(self):
        n = self.n_cells_tot
        ci, cj, ck, le = (np.zeros(n, np.int32) for _ in range(4))
        self.lib.oracle_get_cell_maps(self.h, _p(ci), _p(cj), _p(ck), _p(le))
        return ci, cj, ck, le

    # ---- mc_photon_loop ---------------------------------------------------
    def run(self, n_threads=0, rec=None, n_xI=0, xJ=False, n_Ispec=0, **params):
        r = abi.make_run(**params)
        P = self.P
        t = abi.Tallies(P.n_cells, P.n_lambda, r.struct.N_thet, r.struct.N_phi, xJ=xJ, n_xI=n_xI, n_Ispec=n_Ispec,
                        **abi.grain_tally_sizes(P, r.struct), **_map_args(P, r.struct))
        rec_a = None if rec is None else np.ascontiguousarray(rec, np.float64)
        self._check(self.lib.oracle_run(self.h, r.ref(), t.ref(), int(n_threads), _p(rec_a),
                                        0 if rec_a is None else len(rec_a)))
        return t

    def temp_finale(self):
        T = np.zeros(self.P.n_cells, np.float32)
        self.lib.oracle_temp_finale.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self.lib.oracle_temp_finale(self.h, _p(T)))
        return T

    def temp_finale_of(self, xKJ_abs, xT_ech):
        """Temp_finale (thermal_emission.f90:870-906) of caller-supplied tallies, e.g. the ones downloaded from the GPU"""
        T = np.zeros(self.P.n_cells, np.float32)
        a = np.ascontiguousarray(xKJ_abs, np.float64); b = np.ascontiguousarray(xT_ech, np.int32)
        self.lib.oracle_temp_finale_of.argtypes = [C.c_void_p] * 4
        self._check(self.lib.oracle_temp_finale_of(self.h, _p(a), _p(b), _p(T)))
        return T

    def temp_finale_nlte(self):
        P = self.P
        T = np.zeros((P.grain_RE_nLTE_end - P.grain_RE_nLTE_start + 1, P.n_cells), np.float32, order="F")
        self.lib.oracle_temp_finale_nlte.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self.lib.oracle_temp_finale_nlte(self.h, _p(T)))
        return T

    # ---- deterministic sub-kernels ----------------------------------------
    @staticmethod
    def _f64(*arrs):
        return [np.ascontiguousarray(a, np.float64).copy() for a in arrs]

    def cross_cell(self, x0, y0, z0, u, v, w, icell, previous_cell=None):
        x0, y0, z0, u, v, w = self._f64(x0, y0, z0, u, v, w)
        n = len(x0)
        icell = np.ascontiguousarray(icell, np.int32)
        prev = np.zeros(n, np.int32) if previous_cell is None else np.ascontiguousarray(previous_cell, np.int32)
        x1, y1, z1, l, lc, lv = (np.zeros(n) for _ in range(6))
        nxt = np.zeros(n, np.int32)
        self._check(self.lib.oracle_cross_cell(self.h, C.c_int64(n), _p(x0), _p(y0), _p(z0), _p(u), _p(v), _p(w),
                                               _p(icell), _p(prev), _p(x1), _p(y1), _p(z1), _p(nxt), _p(l), _p(lc), _p(lv)))
        return dict(x1=x1, y1=y1, z1=z1, next_cell=nxt, l=l, l_contrib=lc, l_void_before=lv)

    def index_cell(self, x, y, z):
        x, y, z = self._f64(x, y, z)
        ic = np.zeros(len(x), np.int32)
        self._check(self.lib.oracle_index_cell(self.h, C.c_int64(len(x)), _p(x), _p(y), _p(z), _p(ic)))
        return ic

    def move_to_grid(self, x, y, z, u, v, w):
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        ic = np.zeros(n, np.int32); li = np.zeros(n, np.int32)
        self._check(self.lib.oracle_move_to_grid(self.h, C.c_int64(n), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(ic), _p(li)))
        return dict(x=x, y=y, z=z, icell=ic, lintersect=li)

    def optical_length_tot(self, lam, x, y, z, u, v, w, icell):
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32)
        tau, lmin, lmax = (np.zeros(n) for _ in range(3))
        ns = np.zeros(n, np.int32)
        self._check(self.lib.oracle_optical_length_tot(self.h, C.c_int64(n), C.c_int32(lam), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w),
                                                       _p(icell), _p(tau), _p(lmin), _p(lmax), _p(ns)))
        return dict(tau_tot=tau, lmin=lmin, lmax=lmax, n_steps=ns)

    def define_dark_zone(self, lam, tau_max, r_grid, z_grid, regions=(), dust_sum=None, zj_sup=None, zj_inf=None):
        """define_dark_zone (optical_depth.f90:1425-1651); the result also becomes the oracle's dark zone"""
        P = self.P
        n_az = max(1, P.n_az)
        rg, zg = self._f64(r_grid, z_grid)
        imin = np.ascontiguousarray([r[0] for r in regions], np.int32); imax = np.ascontiguousarray([r[1] for r in regions], np.int32)
        ds = None if dust_sum is None else np.ascontiguousarray(dust_sum, np.float64)
        dark = np.zeros(P.n_cells, np.int32); ri_in = np.zeros(n_az, np.int32); ri_out = np.zeros(n_az, np.int32)
        zs = np.zeros((P.n_rad, n_az), np.int32, order="F") if zj_sup is None else np.asfortranarray(zj_sup, np.int32)
        zi = np.zeros((P.n_rad, n_az), np.int32, order="F") if zj_inf is None else np.asfortranarray(zj_inf, np.int32)
        flag = np.zeros(1, np.int32)
        self._check(self.lib.oracle_define_dark_zone(self.h, C.c_int32(lam), C.c_float(tau_max), _p(rg), _p(zg), C.c_int32(len(regions)),
                                                     _p(imin), _p(imax), _p(ds), _p(dark), _p(ri_in), _p(ri_out), _p(zs), _p(zi), _p(flag)))
        return dict(l_dark_zone=dark, ri_in=ri_in, ri_out=ri_out, zj_sup=zs, zj_inf=zi, l_is_dark_zone=int(flag[0]))

    def init_reemission(self, tab_lambda, tab_delta_lambda):
        """init_reemission (thermal_emission.f90:404-550, LTE cells): log_Qcool_minus_extra_heating (n_T, p_n_cells), kdB_dT_CDF"""
        P = self.P
        tl, td = self._f64(tab_lambda, tab_delta_lambda)
        logQ = np.zeros((P.n_T, P.p_n_cells), np.float64, order="F"); cdf = np.zeros((P.n_lambda, P.n_T, P.p_n_cells), np.float64, order="F")
        self._check(self.lib.oracle_init_reemission(self.h, _p(tl), _p(td), _p(logQ), _p(cdf)))
        return logQ, cdf

    def init_reemission_grains(self, tab_lambda, tab_delta_lambda, C_abs_norm, k_start, k_end):
        """per-grain tables (thermal_emission.f90:551-618): log_E_em (nk, n_T), E_em (nk, n_T), CDF (n_lambda, nk, n_T)"""
        P = self.P
        tl, td = self._f64(tab_lambda, tab_delta_lambda)
        ca = np.asfortranarray(C_abs_norm, np.float32)
        nk = k_end - k_start + 1
        logE = np.zeros((nk, P.n_T), np.float64, order="F"); Eem = np.zeros((nk, P.n_T), np.float64, order="F")
        cdf = np.zeros((P.n_lambda, nk, P.n_T), np.float64, order="F")
        self._check(self.lib.oracle_init_reemission_grains(self.h, _p(tl), _p(td), _p(ca), C.c_int32(ca.shape[0]), C.c_int32(k_start), C.c_int32(k_end),
                                                           _p(logE), _p(Eem), _p(cdf)))
        return logE, Eem, cdf

    def init_dust_source_fct1(self, lam, iRT, n_RT, photon_energy, J_th, xI, n_type_flux, pola, contrib):
        """init_dust_source_fct1 (dust_ray_tracing.f90:636-708): eps (45, 2, ntf, n_cells) from xI_scatt (45, 2, ntf, n_RT, n_cells)"""
        P = self.P
        J = np.ascontiguousarray(J_th, np.float64)
        xi = np.ascontiguousarray(np.asarray(xI, np.float32).reshape(-1))
        n_az, n_th = (1, 1) if P.l3D else (45, 2)
        eps = np.zeros((45, 2, n_type_flux, P.n_cells), np.float64, order="F")
        self._check(self.lib.oracle_init_dust_source_fct1(self.h, C.c_int32(lam), C.c_int32(iRT), C.c_int32(n_RT), C.c_double(photon_energy), _p(J), _p(xi),
                                                          C.c_int32(45), C.c_int32(2), C.c_int32(n_az), C.c_int32(n_th), C.c_int32(n_type_flux),
                                                          C.c_int32(4 if pola else 1), C.c_int32(int(pola)), C.c_int32(int(contrib)), _p(eps)))
        return eps

    def integ_ray_dust(self, lam, x, y, z, u, v, w, icell, tau_dark_zone_obs, eps):
        """integ_ray_dust (optical_depth.f90:1327-1421) with the method-1 source function eps (45, 2, ntf, n_cells): (ntf, n)"""
        P = self.P
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32)
        e = np.asfortranarray(eps, np.float64)
        ntf = e.shape[2]
        out = np.zeros((ntf, n), np.float64, order="F")
        self._check(self.lib.oracle_integ_ray_dust(self.h, C.c_int32(lam), C.c_int64(n), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(icell),
                                                   C.c_float(tau_dark_zone_obs), _p(e), C.c_int32(45), C.c_int32(2), C.c_int32(1 if P.l3D else 45),
                                                   C.c_int32(ntf), _p(out)))
        return out

    def repartition_energie(self, Tdust, tab_lambda, E_stars, E_ISM=None, weight=None, lambda_first=1, lambda_last=None):
        """repartition_energie (thermal_emission.f90:1771-1949, LTE) for a range of wavelengths; uses the oracle's dark zone"""
        P = self.P
        lambda_last = lambda_last or P.n_lambda
        T = np.ascontiguousarray(Tdust, np.float32); tl = np.ascontiguousarray(tab_lambda, np.float64)
        Es = np.ascontiguousarray(E_stars, np.float64)
        Ei = None if E_ISM is None else np.ascontiguousarray(E_ISM, np.float64)
        wt = None if weight is None else np.ascontiguousarray(weight, np.float64)
        E_disk, fs, fd, wn = (np.zeros(P.n_lambda) for _ in range(4))
        prob = np.zeros((P.n_cells + 1, P.n_lambda), np.float64, order="F")
        self._check(self.lib.oracle_repartition_energie(self.h, C.c_int32(lambda_first), C.c_int32(lambda_last), _p(T), _p(tl), _p(Es), _p(Ei), _p(wt),
                                                        _p(E_disk), _p(fs), _p(fd), _p(wn), _p(prob)))
        return dict(E_disk=E_disk, frac_E_stars=fs, frac_E_disk=fd, weight_norm=wn, prob_E_cell=prob)

    def compute_column(self, lam, cx, cy, cz, factor=None):
        """compute_column (optical_depth.f90:328-415): (n_cells, 4) real, column-major; factor None = optical depth at lam"""
        cx, cy, cz = self._f64(cx, cy, cz)
        n = len(cx)
        f = None if factor is None else np.ascontiguousarray(factor, np.float64)
        col = np.zeros((n, 4), np.float32, order="F")
        self._check(self.lib.oracle_compute_column(self.h, C.c_int32(lam), _p(f), _p(cx), _p(cy), _p(cz), _p(col)))
        return col

    def physical_length(self, lam, x, y, z, u, v, w, icell, tau, dark=None):
        if dark is not None:
            self.set_dark_zone(dark)
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32).copy()
        tau = np.ascontiguousarray(tau, np.float32)
        ltot = np.zeros(n, np.float32); fs = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self._check(self.lib.oracle_physical_length(self.h, C.c_int64(n), C.c_int32(lam), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w),
                                                    _p(icell), _p(tau), _p(ltot), _p(fs), _p(alive)))
        return dict(x=x, y=y, z=z, u=u, v=v, w=w, icell=icell, ltot=ltot, flag_sortie=fs, lpacket_alive=alive)

    def distance_to_closest_wall(self, icell, x, y, z):
        x, y, z = self._f64(x, y, z)
        icell = np.ascontiguousarray(icell, np.int32)
        s = np.zeros(len(x))
        self._check(self.lib.oracle_distance_to_closest_wall(self.h, C.c_int64(len(x)), _p(icell), _p(x), _p(y), _p(z), _p(s)))
        return s

    def mrw_tables(self):
        shp = (self.P.n_T, self.P.p_n_cells)
        A, B, Cc = (np.zeros(shp, np.float64, order="F") for _ in range(3))
        self._check(self.lib.oracle_mrw_tables(self.h, _p(A), _p(B), _p(Cc)))
        return A, B, Cc

    def zeta_table(self):
        z = np.zeros(10000)
        self.lib.oracle_zeta_table.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib.oracle_zeta_table(self.h, _p(z), 10000)
        return z

    def sample_zeta(self, zr):
        self.lib.oracle_sample_zeta.argtypes = [C.c_void_p, C.c_double]
        self.lib.oracle_sample_zeta.restype = C.c_double
        return float(self.lib.oracle_sample_zeta(self.h, float(zr)))

    def dark_zone_walker(self):
        """Callable for synthetic.define_dark_zone (step 4 ray walk)."""
        def walk(lam, x, y, z, u, v, w, icell, tau, dark):
            return self.physical_length(lam, x, y, z, u, v, w, icell, tau, dark)["flag_sortie"].astype(bool)
        return walk


def philox(ctr, key):
    lib = _load("liboracle.so")
    c = np.asarray(ctr, np.uint32); k = np.asarray(key, np.uint32); o = np.zeros(4, np.uint32)
    lib.oracle_philox(_p(c), _p(k), _p(o))
    return o


def rng_stream(seed, call_index, packet, n):
    lib = _load("liboracle.so")
    o = np.zeros(n)
    lib.oracle_rng_stream(C.c_uint64(seed), C.c_uint32(call_index), C.c_uint64(packet), C.c_int(n), _p(o))
    return o
