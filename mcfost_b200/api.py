"""Host-side mirror of the reference interface for the photon-packet path.

``PhotonLoop.mc_photon_loop(lambda_in, p_lambda_in, n_photons2, n_phot_lim,
nnfot1_start, laffichage)`` has the reference's own argument list
(src/dust_transfer.f90:439-454); the module-level state the Fortran routine
reads (grid, opacity and emission tables, mode flags) is the ``Problem`` object
plus keyword flags.  Everything is executed by the CUDA library through the C
ABI of include/mcfost_b200.h -- there is no CPU fallback: if the shared library
or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCFOST_B200_LIB") or os.path.join(_HERE, "_lib", "libmcfost_b200.so")   # env override: kernel-tuning builds

# every symbol include/mcfost_b200.h declares
EXPORTS = (
    "mcfost_b200_init", "mcfost_b200_finalize", "mcfost_b200_last_error",
    "mcfost_b200_upload_grid", "mcfost_b200_upload_dark_zone", "mcfost_b200_upload_opacity",
    "mcfost_b200_upload_emission", "mcfost_b200_upload_grains", "mcfost_b200_run", "mcfost_b200_launch", "mcfost_b200_sync",
    "mcfost_b200_tally_buffers", "mcfost_b200_download", "mcfost_b200_last_kernel_ms", "mcfost_b200_stream",
    "mcfost_b200_debug_counters", "mcfost_b200_set_overlap", "mcfost_b200_temp_finale", "mcfost_b200_temp_finale_nlte",
    "mcfost_b200_cross_cell", "mcfost_b200_index_cell", "mcfost_b200_move_to_grid",
    "mcfost_b200_optical_length_tot", "mcfost_b200_physical_length", "mcfost_b200_compute_column", "mcfost_b200_define_dark_zone",
    "mcfost_b200_init_reemission", "mcfost_b200_init_reemission_grains",
    "mcfost_b200_init_dust_source_fct1", "mcfost_b200_integ_ray_dust", "mcfost_b200_repartition_energie",
    "mcfost_b200_distance_to_closest_wall", "mcfost_b200_mrw_tables",
    "mcfost_b200_multi_init", "mcfost_b200_multi_finalize", "mcfost_b200_multi_last_error", "mcfost_b200_multi_n_gpus",
    "mcfost_b200_multi_handle", "mcfost_b200_multi_upload_grid", "mcfost_b200_multi_upload_dark_zone",
    "mcfost_b200_multi_upload_opacity", "mcfost_b200_multi_upload_emission", "mcfost_b200_multi_upload_grains",
    "mcfost_b200_multi_run", "mcfost_b200_multi_temp_finale",
)


def _map_args(P, r):
    kw = abi.map_tally_args(r)
    if "n_xN" in kw and kw["n_xN"] is None:
        kw["n_xN"] = P.n_lambda
    return kw


class McfostB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mcfost_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load_library():
    """dlopen the CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise McfostB200Error(-1, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        lib = C.CDLL(LIB_PATH)
        lib.mcfost_b200_last_error.restype = C.c_char_p
        lib.mcfost_b200_last_error.argtypes = [C.c_void_p]
        lib.mcfost_b200_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        lib.mcfost_b200_finalize.argtypes = [C.c_void_p]
        lib.mcfost_b200_finalize.restype = None
        for fn in ("mcfost_b200_upload_grid", "mcfost_b200_upload_dark_zone", "mcfost_b200_upload_opacity",
                   "mcfost_b200_upload_emission", "mcfost_b200_upload_grains", "mcfost_b200_launch"):
            if hasattr(lib, fn):       # an older build given through MCFOST_B200_LIB (A/B timing) may lack the newest entry points
                getattr(lib, fn).argtypes = [C.c_void_p, C.c_void_p]
        lib.mcfost_b200_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mcfost_b200_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.mcfost_b200_sync.argtypes = [C.c_void_p]
        if hasattr(lib, "mcfost_b200_set_overlap"):
            lib.mcfost_b200_set_overlap.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.mcfost_b200_last_kernel_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        lib.mcfost_b200_stream.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lib.mcfost_b200_tally_buffers.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        _lib = lib
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _DeviceArray:
    """__cuda_array_interface__ view of a library-owned device buffer (so that
    torch.as_tensor(..., device='cuda') can all-reduce it in place with NCCL)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PhotonLoop:
    """One CUDA device running the photon-packet loop for one model."""

    def __init__(self, P, device=0, rank=0, n_ranks=1):
        self.lib = load_library()
        self.P = P
        self.rank, self.n_ranks = int(rank), int(n_ranks)
        self.h = C.c_void_p()
        rc = self.lib.mcfost_b200_init(int(device), C.byref(self.h))
        if rc != 0:
            raise McfostB200Error(rc, self.lib.mcfost_b200_last_error(None).decode())
        self._g = abi.make_grid(P)
        self._check(self.lib.mcfost_b200_upload_grid(self.h, self._g.ref()))
        self._o = abi.make_opacity(P)
        self._check(self.lib.mcfost_b200_upload_opacity(self.h, self._o.ref()))
        self.upload_dark_zone(getattr(P, "l_dark_zone", None))
        self._e = None
        if hasattr(P, "prob_E_cell"):
            self.upload_emission(P)
        self._gr = None
        if hasattr(P, "n_grains_tot"):
            self.upload_grains(P)
        self._last_run = None

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.mcfost_b200_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise McfostB200Error(rc, self.lib.mcfost_b200_last_error(self.h).decode())

    # ---- uploads ------------------------------------------------------------
    def upload_dark_zone(self, dz):
        a = None if dz is None else np.ascontiguousarray(dz, np.int32)
        self._check(self.lib.mcfost_b200_upload_dark_zone(self.h, _p(a)))

    def upload_emission(self, P):
        self._e = abi.make_emission(P)
        self._check(self.lib.mcfost_b200_upload_emission(self.h, self._e.ref()))

    def upload_grains(self, P):
        """Per-grain tables: scattering method 1 and the nLTE / qRE re-emission branches."""
        self._gr = abi.make_grains(P)
        self._check(self.lib.mcfost_b200_upload_grains(self.h, self._gr.ref()))

    # ---- the drop-in --------------------------------------------------------
    def _params(self, lambda_in, p_lambda_in, n_photons2, n_phot_lim, nnfot1_start, laffichage, flags):
        flags = dict(flags)
        flags.setdefault("rank", self.rank)
        flags.setdefault("n_ranks", self.n_ranks)
        flags.setdefault("n_photons_loop", getattr(self.P, "n_photons_loop", 128))
        return abi.make_run(lambda_in=lambda_in, p_lambda_in=p_lambda_in, n_photons2=n_photons2,
                            n_phot_lim=n_phot_lim, nnfot1_start=nnfot1_start, laffichage=int(laffichage), **flags)

    def _tallies(self, r, want_xI=True):
        P = self.P
        xJ = bool(r.struct.lxJ_abs_step1 if r.struct.letape_th else r.struct.lxJ_abs)
        n_xI = 0
        if want_xI and (not r.struct.letape_th) and r.struct.lscatt_ray_tracing1:
            ntf = (4 if r.struct.lsepar_pola else 1) + (4 if r.struct.lsepar_contrib else 0)
            n_xI = abi.N_AZ_RT * 2 * ntf * r.struct.RT_n_incl * r.struct.RT_n_az * P.n_cells
        n_Is = 0
        if (not r.struct.letape_th) and r.struct.lscatt_ray_tracing2 and not r.struct.lscatt_ray_tracing1:
            ntf = (4 if r.struct.lsepar_pola else 1) + (4 if r.struct.lsepar_contrib else 0)
            n_Is = ntf * r.struct.n_theta_I * r.struct.n_phi_I * P.n_cells
        return abi.Tallies(P.n_cells, P.n_lambda, r.struct.N_thet, r.struct.N_phi, xJ=xJ, n_xI=n_xI, n_Ispec=n_Is,
                           **abi.grain_tally_sizes(P, r.struct), **_map_args(P, r.struct))

    def mc_photon_loop(self, lambda_in=1, p_lambda_in=1, n_photons2=1000, n_phot_lim=1.0e30, nnfot1_start=1,
                       laffichage=False, **flags):
        """Blocking call with the reference's argument list; returns the tallies
        (host numpy arrays shaped like the reference's, without the nb_proc dim)."""
        r = self._params(lambda_in, p_lambda_in, n_photons2, n_phot_lim, nnfot1_start, laffichage, flags)
        t = self._tallies(r)
        self._check(self.lib.mcfost_b200_run(self.h, r.ref(), t.ref()))
        self._last_run = r
        return t

    # ---- split form (bench / multi-GPU) ------------------------------------
    def launch(self, lambda_in=1, p_lambda_in=1, n_photons2=1000, n_phot_lim=1.0e30, nnfot1_start=1, **flags):
        r = self._params(lambda_in, p_lambda_in, n_photons2, n_phot_lim, nnfot1_start, False, flags)
        self._check(self.lib.mcfost_b200_launch(self.h, r.ref()))
        self._last_run = r
        return r

    def sync(self):
        self._check(self.lib.mcfost_b200_sync(self.h))

    def set_overlap(self, n_sms_reserved, n_sms_straggler=0):
        """Reserve SMs for the straggler launches so that calls on several handles overlap (0 = off)."""
        self._check(self.lib.mcfost_b200_set_overlap(self.h, int(n_sms_reserved), int(n_sms_straggler)))

    def download(self, r=None, want_xI=True):
        r = r or self._last_run
        t = self._tallies(r, want_xI)
        self._check(self.lib.mcfost_b200_download(self.h, r.ref(), t.ref()))
        return t

    def temp_finale(self):
        """Tdust(n_cells) from the device-resident tallies of the last call (Temp_finale)."""
        T = np.zeros(self.P.n_cells, np.float32)
        self.lib.mcfost_b200_temp_finale.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self.lib.mcfost_b200_temp_finale(self.h, _p(T)))
        return T

    def temp_finale_nlte(self):
        """Tdust_1grain(nLTE grains, n_cells) from the device-resident xJ_abs of the last call (Temp_finale_nLTE)."""
        P = self.P
        T = np.zeros((P.grain_RE_nLTE_end - P.grain_RE_nLTE_start + 1, P.n_cells), np.float32, order="F")
        self.lib.mcfost_b200_temp_finale_nlte.argtypes = [C.c_void_p, C.c_void_p]
        self._check(self.lib.mcfost_b200_temp_finale_nlte(self.h, _p(T)))
        return T

    def last_kernel_ms(self):
        ms = C.c_float()
        self._check(self.lib.mcfost_b200_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)

    def debug_counters(self):
        """Scheduling diagnostics of the last launch (see include/mcfost_b200.h)."""
        out = (C.c_double * 16)()
        self.lib.mcfost_b200_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        self._check(self.lib.mcfost_b200_debug_counters(self.h, out))
        v = list(out)
        names = ("EMIT", "ABSORB", "SCATTER", "FLY")
        return {"steady_ms": v[0], "kernel_ms": v[1],
                "chunk_fill": {n: (v[6 + i] / v[2 + i] if v[2 + i] else 0.0) for i, n in enumerate(names)},
                "visits": {n: v[2 + i] for i, n in enumerate(names)}, "parked": v[10],
                "main_end_ms": v[11], "straggler_end_ms": v[12], "t0_ms": v[13], "straggler_start_ms": v[14], "launches": int(v[15])}

    def stream(self):
        s = C.c_uint64()
        self._check(self.lib.mcfost_b200_stream(self.h, C.byref(s)))
        return int(s.value)

    def tally_buffers(self):
        """(fp64 device view, fp32 device view or None) of the packed tally buffers."""
        p64, n64, p32, n32 = C.c_void_p(), C.c_int64(), C.c_void_p(), C.c_int64()
        self._check(self.lib.mcfost_b200_tally_buffers(self.h, C.byref(p64), C.byref(n64), C.byref(p32), C.byref(n32)))
        a64 = _DeviceArray(p64.value, n64.value, "<f8")
        a32 = _DeviceArray(p32.value, n32.value, "<f4") if n32.value else None
        return a64, a32

    # ---- deterministic sub-kernels -----------------------------------------
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
        self._check(self.lib.mcfost_b200_cross_cell(self.h, C.c_int64(n), _p(x0), _p(y0), _p(z0), _p(u), _p(v), _p(w),
                                                    _p(icell), _p(prev), _p(x1), _p(y1), _p(z1), _p(nxt), _p(l), _p(lc), _p(lv)))
        return dict(x1=x1, y1=y1, z1=z1, next_cell=nxt, l=l, l_contrib=lc, l_void_before=lv)

    def index_cell(self, x, y, z):
        x, y, z = self._f64(x, y, z)
        ic = np.zeros(len(x), np.int32)
        self._check(self.lib.mcfost_b200_index_cell(self.h, C.c_int64(len(x)), _p(x), _p(y), _p(z), _p(ic)))
        return ic

    def move_to_grid(self, x, y, z, u, v, w):
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        ic = np.zeros(n, np.int32); li = np.zeros(n, np.int32)
        self._check(self.lib.mcfost_b200_move_to_grid(self.h, C.c_int64(n), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(ic), _p(li)))
        return dict(x=x, y=y, z=z, icell=ic, lintersect=li)

    def optical_length_tot(self, lam, x, y, z, u, v, w, icell):
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32)
        tau, lmin, lmax = (np.zeros(n) for _ in range(3))
        ns = np.zeros(n, np.int32)
        self._check(self.lib.mcfost_b200_optical_length_tot(self.h, C.c_int64(n), C.c_int32(lam), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w),
                                                            _p(icell), _p(tau), _p(lmin), _p(lmax), _p(ns)))
        return dict(tau_tot=tau, lmin=lmin, lmax=lmax, n_steps=ns)

    def define_dark_zone(self, lam, tau_max, r_grid, z_grid, regions=(), dust_sum=None, zj_sup=None, zj_inf=None):
        """define_dark_zone (optical_depth.f90:1425-1651); the result also becomes this handle's dark zone"""
        P = self.P
        n_az = max(1, P.n_az)
        rg, zg = self._f64(r_grid, z_grid)
        imin = np.ascontiguousarray([r[0] for r in regions], np.int32); imax = np.ascontiguousarray([r[1] for r in regions], np.int32)
        ds = None if dust_sum is None else np.ascontiguousarray(dust_sum, np.float64)
        dark = np.zeros(P.n_cells, np.int32); ri_in = np.zeros(n_az, np.int32); ri_out = np.zeros(n_az, np.int32)
        zs = np.zeros((P.n_rad, n_az), np.int32, order="F") if zj_sup is None else np.asfortranarray(zj_sup, np.int32)
        zi = np.zeros((P.n_rad, n_az), np.int32, order="F") if zj_inf is None else np.asfortranarray(zj_inf, np.int32)
        flag = np.zeros(1, np.int32)
        self._check(self.lib.mcfost_b200_define_dark_zone(self.h, C.c_int32(lam), C.c_float(tau_max), _p(rg), _p(zg), C.c_int32(len(regions)),
                                                          _p(imin) if len(regions) else None, _p(imax) if len(regions) else None, _p(ds),
                                                          _p(dark), _p(ri_in), _p(ri_out), _p(zs), _p(zi), _p(flag)))
        return dict(l_dark_zone=dark, ri_in=ri_in, ri_out=ri_out, zj_sup=zs, zj_inf=zi, l_is_dark_zone=int(flag[0]))

    def init_reemission(self, tab_lambda, tab_delta_lambda, download=True):
        """init_reemission (thermal_emission.f90:404-550) on the device; the tables become the handle's thermal tables"""
        P = self.P
        tl, td = self._f64(tab_lambda, tab_delta_lambda)
        logQ = cdf = None
        if download:
            logQ = np.zeros((P.n_T, P.p_n_cells), np.float64, order="F"); cdf = np.zeros((P.n_lambda, P.n_T, P.p_n_cells), np.float64, order="F")
        self._check(self.lib.mcfost_b200_init_reemission(self.h, _p(tl), _p(td), _p(logQ), _p(cdf)))
        return logQ, cdf

    def init_reemission_grains(self, tab_lambda, tab_delta_lambda, C_abs_norm, k_start, k_end):
        """per-grain tables of init_reemission (thermal_emission.f90:551-618): log_E_em (nk, n_T), E_em (nk, n_T), CDF (n_lambda, nk, n_T)"""
        P = self.P
        tl, td = self._f64(tab_lambda, tab_delta_lambda)
        ca = np.asfortranarray(C_abs_norm, np.float32)
        nk = k_end - k_start + 1
        logE = np.zeros((nk, P.n_T), np.float64, order="F"); Eem = np.zeros((nk, P.n_T), np.float64, order="F")
        cdf = np.zeros((P.n_lambda, nk, P.n_T), np.float64, order="F")
        self._check(self.lib.mcfost_b200_init_reemission_grains(self.h, _p(tl), _p(td), _p(ca), C.c_int32(ca.shape[0]), C.c_int32(k_start),
                                                                C.c_int32(k_end), _p(logE), _p(Eem), _p(cdf)))
        return logE, Eem, cdf

    def init_dust_source_fct1(self, lam, iRT, photon_energy, J_th, n_type_flux, download=True):
        """init_dust_source_fct1 (dust_ray_tracing.f90:636-708) from the xI_scatt tally on the device; eps (45, 2, ntf, n_cells)"""
        J = np.ascontiguousarray(J_th, np.float64)
        eps = np.zeros((45, 2, n_type_flux, self.P.n_cells), np.float64, order="F") if download else None
        self._check(self.lib.mcfost_b200_init_dust_source_fct1(self.h, C.c_int32(lam), C.c_int32(iRT), C.c_double(photon_energy), _p(J), _p(eps)))
        return eps

    def integ_ray_dust(self, lam, x, y, z, u, v, w, icell, tau_dark_zone_obs, n_type_flux):
        """integ_ray_dust (optical_depth.f90:1327-1421), method-1 source function: (n_type_flux, n)"""
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32)
        out = np.zeros((n_type_flux, n), np.float64, order="F")
        self._check(self.lib.mcfost_b200_integ_ray_dust(self.h, C.c_int32(lam), C.c_int64(n), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w), _p(icell),
                                                        C.c_float(tau_dark_zone_obs), _p(out)))
        return out

    def repartition_energie(self, Tdust, tab_lambda, E_stars, E_ISM=None, weight=None, lambda_first=1, lambda_last=None, download=True):
        """repartition_energie (thermal_emission.f90:1771-1949, LTE) on the device; the tables become the handle's emission tables"""
        P = self.P
        lambda_last = lambda_last or P.n_lambda
        T = np.ascontiguousarray(Tdust, np.float32); tl = np.ascontiguousarray(tab_lambda, np.float64)
        Es = np.ascontiguousarray(E_stars, np.float64)
        Ei = None if E_ISM is None else np.ascontiguousarray(E_ISM, np.float64)
        wt = None if weight is None else np.ascontiguousarray(weight, np.float64)
        E_disk, fs, fd, wn = (np.zeros(P.n_lambda) for _ in range(4))
        prob = np.zeros((P.n_cells + 1, P.n_lambda), np.float64, order="F") if download else None
        self._check(self.lib.mcfost_b200_repartition_energie(self.h, C.c_int32(lambda_first), C.c_int32(lambda_last), _p(T), _p(tl), _p(Es), _p(Ei),
                                                             _p(wt), _p(E_disk), _p(fs), _p(fd), _p(wn), _p(prob)))
        return dict(E_disk=E_disk, frac_E_stars=fs, frac_E_disk=fd, weight_norm=wn, prob_E_cell=prob)

    def compute_column(self, lam, cx, cy, cz, factor=None):
        """compute_column (optical_depth.f90:328-415): (n_cells, 4) real, column-major; factor None = optical depth at lam"""
        cx, cy, cz = self._f64(cx, cy, cz)
        f = None if factor is None else np.ascontiguousarray(factor, np.float64)
        col = np.zeros((self.P.n_cells, 4), np.float32, order="F")
        self._check(self.lib.mcfost_b200_compute_column(self.h, C.c_int32(lam), _p(f), _p(cx), _p(cy), _p(cz), _p(col)))
        return col

    def physical_length(self, lam, x, y, z, u, v, w, icell, tau, dark=None):
        if dark is not None:
            self.upload_dark_zone(dark)
        x, y, z, u, v, w = self._f64(x, y, z, u, v, w)
        n = len(x)
        icell = np.ascontiguousarray(icell, np.int32).copy()
        tau = np.ascontiguousarray(tau, np.float32)
        ltot = np.zeros(n, np.float32); fs = np.zeros(n, np.int32); alive = np.zeros(n, np.int32)
        self._check(self.lib.mcfost_b200_physical_length(self.h, C.c_int64(n), C.c_int32(lam), _p(x), _p(y), _p(z), _p(u), _p(v), _p(w),
                                                         _p(icell), _p(tau), _p(ltot), _p(fs), _p(alive)))
        return dict(x=x, y=y, z=z, u=u, v=v, w=w, icell=icell, ltot=ltot, flag_sortie=fs, lpacket_alive=alive)

    def distance_to_closest_wall(self, icell, x, y, z):
        """distance_to_closest_wall (grid.f90 procedure pointer) for points inside real cells."""
        x, y, z = self._f64(x, y, z)
        icell = np.ascontiguousarray(icell, np.int32)
        s = np.zeros(len(x))
        self._check(self.lib.mcfost_b200_distance_to_closest_wall(self.h, C.c_int64(len(x)), _p(icell), _p(x), _p(y), _p(z), _p(s)))
        return s

    def mrw_tables(self):
        """Mean opacities of the modified random walk, (n_T, p_n_cells) each: A, B, C (see include/mcfost_b200.h)."""
        shp = (self.P.n_T, self.P.p_n_cells)
        A, B, Cc = (np.zeros(shp, np.float64, order="F") for _ in range(3))
        self._check(self.lib.mcfost_b200_mrw_tables(self.h, _p(A), _p(B), _p(Cc)))
        return A, B, Cc

    def dark_zone_walker(self):
        """Step-4 ray walk of define_dark_zone (optical_depth.f90:1519-1550) on the GPU."""
        def walk(lam, x, y, z, u, v, w, icell, tau, dark):
            return self.physical_length(lam, x, y, z, u, v, w, icell, tau, dark)["flag_sortie"].astype(bool)
        return walk


class MultiPhotonLoop:
    """The n GPUs of one node behind ONE object and one call (mcfost_b200_multi_*): grid and tables replicated,
    chunks dealt round-robin, every tally merged by NCCL inside the library before the call returns."""

    def __init__(self, P, n_gpus, devices=None):
        self.lib = load_library()
        L = self.lib
        L.mcfost_b200_multi_init.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.mcfost_b200_multi_finalize.argtypes = [C.c_void_p]; L.mcfost_b200_multi_finalize.restype = None
        L.mcfost_b200_multi_last_error.argtypes = [C.c_void_p]; L.mcfost_b200_multi_last_error.restype = C.c_char_p
        L.mcfost_b200_multi_handle.argtypes = [C.c_void_p, C.c_int]; L.mcfost_b200_multi_handle.restype = C.c_void_p
        for fn in ("upload_grid", "upload_dark_zone", "upload_opacity", "upload_emission", "upload_grains", "temp_finale"):
            getattr(L, "mcfost_b200_multi_" + fn).argtypes = [C.c_void_p, C.c_void_p]
        L.mcfost_b200_multi_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        self.P, self.n_gpus = P, int(n_gpus)
        self.m = C.c_void_p()
        dev = None if devices is None else np.ascontiguousarray(devices, np.int32)
        rc = L.mcfost_b200_multi_init(self.n_gpus, _p(dev), C.byref(self.m))
        if rc != 0:
            raise McfostB200Error(rc, "multi_init failed (no such devices, or libnccl.so.2 not found): " + L.mcfost_b200_last_error(None).decode())
        self._g = abi.make_grid(P); self._check(L.mcfost_b200_multi_upload_grid(self.m, self._g.ref()))
        self._o = abi.make_opacity(P); self._check(L.mcfost_b200_multi_upload_opacity(self.m, self._o.ref()))
        self.upload_dark_zone(getattr(P, "l_dark_zone", None))
        self._e = None
        if hasattr(P, "prob_E_cell"):
            self.upload_emission(P)
        self._gr = None
        if hasattr(P, "n_grains_tot"):
            self._gr = abi.make_grains(P); self._check(L.mcfost_b200_multi_upload_grains(self.m, self._gr.ref()))

    def _check(self, rc):
        if rc != 0:
            raise McfostB200Error(rc, self.lib.mcfost_b200_multi_last_error(self.m).decode())

    def upload_dark_zone(self, dz):
        a = None if dz is None else np.ascontiguousarray(dz, np.int32)
        self._check(self.lib.mcfost_b200_multi_upload_dark_zone(self.m, _p(a)))

    def upload_emission(self, P):
        self._e = abi.make_emission(P)
        self._check(self.lib.mcfost_b200_multi_upload_emission(self.m, self._e.ref()))

    def mc_photon_loop(self, lambda_in=1, p_lambda_in=1, n_photons2=1000, n_phot_lim=1.0e30, nnfot1_start=1, laffichage=False, **flags):
        flags = dict(flags)
        flags.setdefault("n_photons_loop", getattr(self.P, "n_photons_loop", 128))
        r = abi.make_run(lambda_in=lambda_in, p_lambda_in=p_lambda_in, n_photons2=n_photons2, n_phot_lim=n_phot_lim,
                         nnfot1_start=nnfot1_start, laffichage=int(laffichage), **flags)
        t = PhotonLoop._tallies(self, r)
        self._check(self.lib.mcfost_b200_multi_run(self.m, r.ref(), t.ref()))
        return t

    def temp_finale(self):
        T = np.zeros(self.P.n_cells, np.float32)
        self._check(self.lib.mcfost_b200_multi_temp_finale(self.m, _p(T)))
        return T

    def close(self):
        if getattr(self, "m", None) is not None and self.m:
            self.lib.mcfost_b200_multi_finalize(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
