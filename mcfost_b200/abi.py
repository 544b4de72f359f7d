"""ctypes mirror of include/mcfost_b200.h (the C ABI that replaces
``mc_photon_loop``, reference src/dust_transfer.f90:439-572).

Only data layout lives here: the structs are filled from numpy arrays that are
kept alive by the returned holder objects.  Arrays are Fortran column-major and
cell / wavelength indices 1-based exactly as the reference's module variables.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)

MCB_OK = 0
MCB_ERR_NO_DEVICE = 1
MCB_ERR_BAD_ARG = 2
MCB_ERR_CELL_MAP = 3
MCB_ERR_CUDA = 4
MCB_ERR_UNSUPPORTED = 5
MCB_ERR_STATE = 6

MCB_GRID_CYL, MCB_GRID_SPH, MCB_GRID_VORONOI = 1, 2, 3
NANG_SCATT = 180
N_AZ_RT = 45


class mcb_grid(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("l3D", C.c_int32),
        ("n_rad", C.c_int32), ("nz", C.c_int32), ("n_az", C.c_int32),
        ("n_cells", C.c_int32),
        ("Rmax2", C.c_double), ("zmaxmax", C.c_double),
        ("r_lim", c_double_p), ("r_lim_2", c_double_p), ("r_lim_3", c_double_p),
        ("z_lim", c_double_p), ("zmax", c_double_p),
        ("tan_theta_lim", c_double_p), ("theta_lim", c_double_p),
        ("tan_phi_lim", c_double_p), ("volume", c_double_p),
        ("n_cells_tot", C.c_int32),
        ("cell_map_i", c_int32_p), ("cell_map_j", c_int32_p), ("cell_map_k", c_int32_p),
        ("vor_xyz", c_double_p), ("vor_h", c_double_p),
        ("vor_first", c_int32_p), ("vor_last", c_int32_p),
        ("vor_was_cut", c_int32_p), ("vor_is_star", c_int32_p), ("vor_is_star_neighbour", c_int32_p),
        ("neighbours_list", c_int32_p), ("n_neighbours_tot", C.c_int64),
        ("wall_x", (C.c_float * 4) * 6),
        ("cutting_distance_o_h", C.c_double),
        ("n_stars", C.c_int32),
        ("star_xyzr", c_double_p), ("star_icell", c_int32_p), ("star_out_model", c_int32_p),
        ("w_lim", c_double_p), ("sin_phi_lim", c_double_p), ("cos_phi_lim", c_double_p),
    ]


class mcb_opacity(C.Structure):
    _fields_ = [
        ("n_lambda", C.c_int32), ("p_n_cells", C.c_int32),
        ("p_n_lambda_pos", C.c_int32), ("n_T", C.c_int32),
        ("kappa", c_double_p), ("kappa_abs_LTE", c_double_p), ("kappa_factor", c_double_p),
        ("tab_albedo_pos", c_float_p), ("tab_g_pos", c_float_p),
        ("prob_s11_pos", c_float_p), ("tab_s11_pos", c_float_p),
        ("tab_s12_o_s11_pos", c_float_p), ("tab_s22_o_s11_pos", c_float_p),
        ("tab_s33_o_s11_pos", c_float_p), ("tab_s34_o_s11_pos", c_float_p),
        ("tab_s44_o_s11_pos", c_float_p),
        ("log_Qcool_minus_extra_heating", c_double_p), ("kdB_dT_CDF", c_double_p),
        ("tab_Temp", c_float_p), ("T_min", C.c_float),
    ]


class mcb_emission(C.Structure):
    _fields_ = [
        ("spectre_emission_cumul", c_double_p),
        ("frac_E_stars", c_double_p), ("frac_E_disk", c_double_p),
        ("prob_E_cell", c_double_p), ("CDF_E_star", c_float_p),
        ("L_packet_th", C.c_double), ("E_paquet", C.c_double),
        ("R_ISM", C.c_double), ("centre_ISM", C.c_double * 3),
        ("correct_E_emission", c_double_p),
    ]


_GR_I32 = ("grain_zone", "l_RE")
_GR_F32 = ("C_abs", "C_abs_norm", "C_sca", "tab_g", "prob_s11", "tab_s11", "tab_s12", "tab_s22", "tab_s33", "tab_s34", "tab_s44")


class mcb_grains(C.Structure):
    _fields_ = [
        ("n_grains_tot", C.c_int32), ("n_dens", C.c_int32),
        ("grain_RE_LTE_start", C.c_int32), ("grain_RE_LTE_end", C.c_int32),
        ("grain_RE_nLTE_start", C.c_int32), ("grain_RE_nLTE_end", C.c_int32),
        ("grain_nRE_start", C.c_int32), ("grain_nRE_end", C.c_int32),
        ("grain_zone", c_int32_p), ("n_grains", c_double_p), ("dust_density_o_n_grains", c_double_p),
        ("C_abs", c_float_p), ("C_abs_norm", c_float_p), ("C_sca", c_float_p), ("tab_g", c_float_p),
        ("prob_s11", c_float_p),
        ("tab_s11", c_float_p), ("tab_s12", c_float_p), ("tab_s22", c_float_p),
        ("tab_s33", c_float_p), ("tab_s34", c_float_p), ("tab_s44", c_float_p),
        ("ksca_CDF", c_double_p),
        ("kappa_abs_nLTE", c_double_p), ("kabs_nLTE_CDF", c_double_p),
        ("log_E_em_1grain", c_double_p), ("kdB_dT_1grain_nLTE_CDF", c_double_p),
        ("kappa_abs_RE", c_double_p), ("proba_abs_RE", c_double_p),
        ("Proba_abs_RE_LTE", c_double_p), ("Proba_abs_RE_LTE_p_nLTE", c_double_p),
        ("log_E_em_1grain_nRE", c_double_p), ("kdB_dT_1grain_nRE_CDF", c_double_p),
        ("l_RE", c_int32_p), ("J0", c_double_p), ("kdB_dT_1grain_LTE_CDF", c_double_p),
    ]


class mcb_run_params(C.Structure):
    _fields_ = [
        ("lambda_in", C.c_int32), ("p_lambda_in", C.c_int32), ("n_photons2", C.c_int32),
        ("n_phot_lim", C.c_float), ("nnfot1_start", C.c_int32), ("laffichage", C.c_int32),
        ("n_photons_loop", C.c_int32),
        ("letape_th", C.c_int32), ("lmono", C.c_int32), ("lmono0", C.c_int32),
        ("lscatt_ray_tracing1", C.c_int32), ("lscatt_ray_tracing2", C.c_int32),
        ("lsepar_pola", C.c_int32), ("lsepar_contrib", C.c_int32),
        ("lscattering_method1", C.c_int32), ("lmethod_aniso1", C.c_int32),
        ("lisotropic", C.c_int32),
        ("l_sym_centrale", C.c_int32), ("l_sym_axiale", C.c_int32),
        ("lonly_LTE", C.c_int32), ("lxJ_abs_step1", C.c_int32), ("lxJ_abs", C.c_int32),
        ("N_thet", C.c_int32), ("N_phi", C.c_int32), ("capt_sup", C.c_int32),
        ("RT_n_incl", C.c_int32), ("RT_n_az", C.c_int32),
        ("tab_u_rt", c_double_p), ("tab_v_rt", c_double_p), ("tab_w_rt", c_double_p),
        ("seed", C.c_uint64), ("call_index", C.c_uint32),
        ("rank", C.c_int32), ("n_ranks", C.c_int32), ("reset_tallies", C.c_int32),
        ("loutput_mc", C.c_int32), ("n_theta_I", C.c_int32), ("n_phi_I", C.c_int32),
        ("lonly_nLTE", C.c_int32), ("lRE_nLTE", C.c_int32), ("lnRE", C.c_int32),
        ("low_mem_th_emission_nLTE", C.c_int32), ("low_mem_scattering", C.c_int32),
        ("npix_x", C.c_int32), ("npix_y", C.c_int32), ("zoom", C.c_float), ("map_size", C.c_double),
        ("cos_disk", C.c_double), ("sin_disk", C.c_double), ("l_sym_ima", C.c_int32),
        ("lonly_capt_interet", C.c_int32), ("capt_inf", C.c_int32), ("lorigine", C.c_int32), ("capt_interet", C.c_int32),
        ("low_mem_th_emission", C.c_int32), ("lweight_emission", C.c_int32), ("lspot", C.c_int32),
        ("T_spot", C.c_float), ("surf_fraction_spot", C.c_float), ("theta_spot", C.c_float), ("phi_spot", C.c_float),
        ("star1_T", C.c_double), ("tab_lambda", c_double_p), ("lxN_abs", C.c_int32),
        ("lMRW", C.c_int32), ("gamma_MRW", C.c_float), ("lcount_sent", C.c_int32), ("max_inflight_fraction", C.c_float),
        ("lISM_loop", C.c_int32),
    ]


class mcb_tallies(C.Structure):
    _fields_ = [
        ("xKJ_abs", c_double_p), ("xJ_abs", c_double_p), ("xT_ech", c_int32_p),
        ("n_phot_envoyes", c_double_p),
        ("sed", c_double_p), ("sed_q", c_double_p), ("sed_u", c_double_p), ("sed_v", c_double_p),
        ("n_phot_sed", c_double_p),
        ("sed_star", c_double_p), ("sed_star_scat", c_double_p),
        ("sed_disk", c_double_p), ("sed_disk_scat", c_double_p),
        ("xI_scatt", c_float_p), ("N_type_flux", C.c_int32),
        ("I_spec", c_float_p), ("I_spec_star", c_float_p),
        ("stats", c_double_p),
        ("xT_ech_1grain", c_int32_p), ("xT_ech_1grain_nRE", c_int32_p), ("E_abs_nRE", c_double_p),
        ("stokes_map", c_double_p), ("star_origin", c_double_p), ("disk_origin", c_double_p), ("xN_abs", c_double_p),
    ]


_CT = {np.dtype("float64"): c_double_p, np.dtype("float32"): c_float_p, np.dtype("int32"): c_int32_p}


def ptr(a, dtype):
    """Pointer to a Fortran-contiguous numpy array of exactly ``dtype`` (None -> NULL)."""
    dt = np.dtype(dtype)
    if a is None:
        return C.cast(None, _CT[dt])
    if a.dtype != dt:
        raise TypeError(f"expected {dt}, got {a.dtype}")
    if not (a.flags["F_CONTIGUOUS"] or a.flags["C_CONTIGUOUS"] and a.ndim <= 1):
        raise ValueError("array must be Fortran-contiguous")
    return a.ctypes.data_as(_CT[dt])


def farray(a, dtype):
    """Fortran-contiguous copy/view with the exact dtype the ABI wants."""
    return np.asfortranarray(np.asarray(a, dtype=dtype))


class Holder:
    """A filled ctypes struct plus the numpy arrays that back its pointers."""

    def __init__(self, struct, keep):
        self.struct = struct
        self.keep = keep

    def ref(self):
        return C.byref(self.struct)


def make_grid(P) -> Holder:
    g = mcb_grid()
    keep = {}

    def put(name, dtype):
        a = getattr(P, name, None)
        if a is not None:
            a = farray(a, dtype)
            keep[name] = a
        setattr(g, name, ptr(a, dtype))

    g.kind, g.l3D = int(P.kind), int(P.l3D)
    g.n_rad, g.nz, g.n_az, g.n_cells = int(P.n_rad), int(P.nz), int(P.n_az), int(P.n_cells)
    g.Rmax2, g.zmaxmax = float(P.Rmax2), float(P.zmaxmax)
    for name in ("r_lim", "r_lim_2", "r_lim_3", "z_lim", "zmax", "tan_theta_lim", "theta_lim",
                 "tan_phi_lim", "volume", "vor_xyz", "vor_h", "star_xyzr", "w_lim", "sin_phi_lim", "cos_phi_lim"):
        put(name, np.float64)
    for name in ("cell_map_i", "cell_map_j", "cell_map_k", "vor_first", "vor_last", "vor_was_cut",
                 "vor_is_star", "vor_is_star_neighbour", "neighbours_list", "star_icell", "star_out_model"):
        put(name, np.int32)
    g.n_cells_tot = int(getattr(P, "n_cells_tot", 0) or 0)
    nl = getattr(P, "neighbours_list", None)
    g.n_neighbours_tot = 0 if nl is None else int(len(nl))
    wall = getattr(P, "wall_x", None)
    if wall is not None:
        for i in range(6):
            for j in range(4):
                g.wall_x[i][j] = float(wall[i][j])
    g.cutting_distance_o_h = float(getattr(P, "cutting_distance_o_h", 0.0) or 0.0)
    g.n_stars = int(P.n_stars)
    return Holder(g, keep)


def make_opacity(P) -> Holder:
    o = mcb_opacity()
    keep = {}
    o.n_lambda, o.p_n_cells = int(P.n_lambda), int(P.p_n_cells)
    o.p_n_lambda_pos, o.n_T = int(P.p_n_lambda_pos), int(P.n_T)
    for name in ("kappa", "kappa_abs_LTE", "kappa_factor", "log_Qcool_minus_extra_heating", "kdB_dT_CDF"):
        a = getattr(P, name)
        if a is not None:      # (the two thermal tables may be left to mcfost_b200_init_reemission)
            a = farray(a, np.float64)
            keep[name] = a
        setattr(o, name, ptr(a, np.float64))
    for name in ("tab_albedo_pos", "tab_g_pos", "prob_s11_pos", "tab_s11_pos", "tab_s12_o_s11_pos",
                 "tab_s22_o_s11_pos", "tab_s33_o_s11_pos", "tab_s34_o_s11_pos", "tab_s44_o_s11_pos", "tab_Temp"):
        a = getattr(P, name, None)
        if a is not None:
            a = farray(a, np.float32)
            keep[name] = a
        setattr(o, name, ptr(a, np.float32))
    o.T_min = float(P.T_min)
    return Holder(o, keep)


def make_emission(P) -> Holder:
    e = mcb_emission()
    keep = {}
    for name in ("spectre_emission_cumul", "frac_E_stars", "frac_E_disk", "prob_E_cell"):
        a = getattr(P, name)
        if a is not None:      # (the last three may be left to mcfost_b200_repartition_energie)
            a = farray(a, np.float64)
            keep[name] = a
        setattr(e, name, ptr(a, np.float64))
    a = farray(P.CDF_E_star, np.float32)
    keep["CDF_E_star"] = a
    e.CDF_E_star = ptr(a, np.float32)
    e.L_packet_th, e.E_paquet = float(P.L_packet_th), float(P.E_paquet)
    e.R_ISM = float(getattr(P, "R_ISM", 0.0))
    c = getattr(P, "centre_ISM", (0.0, 0.0, 0.0))
    for i in range(3):
        e.centre_ISM[i] = float(c[i])
    a = getattr(P, "correct_E_emission", None)
    if a is not None:
        a = farray(a, np.float64)
        keep["correct_E_emission"] = a
    e.correct_E_emission = ptr(a, np.float64)
    return Holder(e, keep)


def make_grains(P) -> Holder:
    """Per-grain tables (scattering method 1, nLTE / qRE re-emission); absent attributes -> NULL."""
    g = mcb_grains()
    keep = {}
    for name, ctype in mcb_grains._fields_:
        if ctype is C.c_int32:
            setattr(g, name, int(getattr(P, name, 0) or 0))
            continue
        dt = np.int32 if name in _GR_I32 else np.float32 if name in _GR_F32 else np.float64
        a = getattr(P, name, None)
        if a is not None:
            a = farray(a, dt)
            keep[name] = a
        setattr(g, name, ptr(a, dt))
    return Holder(g, keep)


def make_run(**kw) -> Holder:
    """Run parameters with the reference's defaults for a thermal step
    (dust_transfer.f90:597-617, read_param.f90:145,180-184)."""
    d = dict(lambda_in=1, p_lambda_in=1, n_photons2=1000, n_phot_lim=1.0e30, nnfot1_start=1, laffichage=0,
             n_photons_loop=128, letape_th=1, lmono=0, lmono0=0, lscatt_ray_tracing1=0, lscatt_ray_tracing2=0,
             lsepar_pola=0, lsepar_contrib=0, lscattering_method1=0, lmethod_aniso1=1, lisotropic=0,
             l_sym_centrale=1, l_sym_axiale=1, lonly_LTE=1, lxJ_abs_step1=0, lxJ_abs=0,
             N_thet=10, N_phi=1, capt_sup=2, RT_n_incl=0, RT_n_az=0,
             tab_u_rt=None, tab_v_rt=None, tab_w_rt=None,
             seed=269753, call_index=0, rank=0, n_ranks=1, reset_tallies=1,
             loutput_mc=0, n_theta_I=15, n_phi_I=15,
             lonly_nLTE=0, lRE_nLTE=0, lnRE=0, low_mem_th_emission_nLTE=0, low_mem_scattering=1,
             npix_x=0, npix_y=0, zoom=1.0, map_size=0.0, cos_disk=1.0, sin_disk=0.0, l_sym_ima=0,
             lonly_capt_interet=0, capt_inf=1, lorigine=0, capt_interet=1,
             low_mem_th_emission=0, lweight_emission=0, lspot=0, T_spot=0.0, surf_fraction_spot=0.0, theta_spot=0.0,
             phi_spot=0.0, star1_T=0.0, tab_lambda=None, lxN_abs=0,
             lMRW=0, gamma_MRW=2.0, lcount_sent=0, max_inflight_fraction=0.0, lISM_loop=0)
    unknown = set(kw) - set(d)
    if unknown:
        raise TypeError(f"unknown run parameter(s): {sorted(unknown)}")
    d.update(kw)
    r = mcb_run_params()
    keep = {}
    for k, v in d.items():
        if k in ("tab_u_rt", "tab_v_rt", "tab_w_rt", "tab_lambda"):
            a = None if v is None else farray(v, np.float64)
            keep[k] = a
            setattr(r, k, ptr(a, np.float64))
        else:
            setattr(r, k, v)
    return Holder(r, keep)


class Tallies:
    """Caller-allocated tally arrays (shapes of the reference minus the nb_proc dim)."""

    def __init__(self, n_cells, n_lambda, N_thet=10, N_phi=1, xJ=False, n_xI=0, n_Ispec=0, n_nLTE=0, n_nRE=0,
                 map_shape=None, origin=False, n_xN=0):
        self.xKJ_abs = np.zeros(n_cells, np.float64)
        self.xJ_abs = np.zeros((n_cells, n_lambda), np.float64, order="F") if xJ else None
        self.xT_ech = np.zeros(n_cells, np.int32)
        self.n_phot_envoyes = np.zeros(n_lambda, np.float64)
        shp = (n_lambda, N_thet, N_phi)
        for name in ("sed", "sed_q", "sed_u", "sed_v", "n_phot_sed", "sed_star", "sed_star_scat",
                     "sed_disk", "sed_disk_scat"):
            setattr(self, name, np.zeros(shp, np.float64, order="F"))
        self.xI_scatt = np.zeros(n_xI, np.float32) if n_xI else None
        self.I_spec = np.zeros(n_Ispec, np.float32) if n_Ispec else None
        self.I_spec_star = np.zeros(n_cells, np.float32) if n_Ispec else None
        self.stats = np.zeros(12, np.float64)
        self.xT_ech_1grain = np.zeros((n_nLTE, n_cells), np.int32, order="F") if n_nLTE else None
        self.xT_ech_1grain_nRE = np.zeros((n_nRE, n_cells), np.int32, order="F") if n_nRE else None
        self.E_abs_nRE = np.zeros(1, np.float64)
        self.stokes_map = np.zeros(map_shape, np.float64, order="F") if map_shape else None
        self.star_origin = np.zeros(n_lambda, np.float64) if origin else None
        self.disk_origin = np.zeros((n_lambda, n_cells), np.float64, order="F") if origin else None
        self.xN_abs = np.zeros((n_cells, n_xN), np.float64, order="F") if n_xN else None
        t = mcb_tallies()
        for name in ("xKJ_abs", "xJ_abs", "n_phot_envoyes", "sed", "sed_q", "sed_u", "sed_v", "n_phot_sed",
                     "sed_star", "sed_star_scat", "sed_disk", "sed_disk_scat", "stats"):
            setattr(t, name, ptr(getattr(self, name), np.float64))
        t.xT_ech = ptr(self.xT_ech, np.int32)
        t.xT_ech_1grain = ptr(self.xT_ech_1grain, np.int32)
        t.xT_ech_1grain_nRE = ptr(self.xT_ech_1grain_nRE, np.int32)
        t.E_abs_nRE = ptr(self.E_abs_nRE, np.float64)
        t.stokes_map = ptr(self.stokes_map, np.float64)
        t.star_origin = ptr(self.star_origin, np.float64)
        t.disk_origin = ptr(self.disk_origin, np.float64)
        t.xN_abs = ptr(self.xN_abs, np.float64)
        t.xI_scatt = ptr(self.xI_scatt, np.float32)
        t.I_spec = ptr(self.I_spec, np.float32)
        t.I_spec_star = ptr(self.I_spec_star, np.float32)
        t.N_type_flux = 0
        self.struct = t

    def ref(self):
        return C.byref(self.struct)


def grain_tally_sizes(P, r):
    """Extents of xT_ech_1grain / xT_ech_1grain_nRE for this run (thermal_emission.f90:163,186)."""
    n1 = (P.grain_RE_nLTE_end - P.grain_RE_nLTE_start + 1) if (r.lRE_nLTE and hasattr(P, "grain_RE_nLTE_start")) else 0
    n2 = (P.grain_nRE_end - P.grain_nRE_start + 1) if (r.lnRE and hasattr(P, "grain_nRE_start")) else 0
    return dict(n_nLTE=int(n1), n_nRE=int(n2))


def map_tally_args(r):
    """Tallies(...) keyword arguments for the Monte Carlo photon maps / origin tallies of this run."""
    kw = {}
    if r.lmono0 and r.loutput_mc and r.npix_x > 0 and r.npix_y > 0:
        ntf = (4 if r.lsepar_pola else 1) + (4 if r.lsepar_contrib else 0)
        kw["map_shape"] = (r.npix_x, r.npix_y, r.N_thet, r.N_phi, ntf)
    if r.lorigine:
        kw["origin"] = True
    if r.lxN_abs and (r.letape_th or r.lxJ_abs):
        kw["n_xN"] = 1 if r.letape_th else None      # None -> n_lambda, filled in by the caller
    return kw
