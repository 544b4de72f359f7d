"""Write a synthetic problem as a flat binary file for shim/c_driver.c (records: 32-byte name, int32 dtype code
0 = f64 / 1 = f32 / 2 = i32, int64 count, data in Fortran order).  usage: dump_problem.py out.bin [n2]"""
import sys, os, struct
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mcfost_b200 import synthetic as S

GRID = ("r_lim", "r_lim_2", "r_lim_3", "z_lim", "zmax", "tan_phi_lim", "volume", "star_xyzr", "r_grid", "z_grid", "tab_lambda", "tab_delta_lambda")
GRID_I = ("cell_map_i", "cell_map_j", "cell_map_k", "star_icell", "star_out_model")
OPA = ("kappa", "kappa_abs_LTE", "kappa_factor", "log_Qcool_minus_extra_heating", "kdB_dT_CDF")
OPA_F = ("tab_albedo_pos", "tab_g_pos", "prob_s11_pos", "tab_s11_pos", "tab_s12_o_s11_pos", "tab_s22_o_s11_pos",
         "tab_s33_o_s11_pos", "tab_s34_o_s11_pos", "tab_s44_o_s11_pos", "tab_Temp")
EMI = ("spectre_emission_cumul", "frac_E_stars", "frac_E_disk", "prob_E_cell")


def dump(P, path):
    with open(path, "wb") as f:
        def rec(name, a, code):
            dt = (np.float64, np.float32, np.int32)[code]
            a = np.asfortranarray(np.asarray(a, dtype=dt))
            f.write(name.encode().ljust(32, b"\0")); f.write(struct.pack("<iq", code, a.size)); f.write(a.tobytes(order="F"))
        for k in ("kind", "l3D", "n_rad", "nz", "n_az", "n_cells", "n_cells_tot", "n_stars", "n_lambda", "p_n_cells", "p_n_lambda_pos", "n_T", "lambda_seuil"):
            rec(k, [int(getattr(P, k))], 2)
        for k in ("Rmax2", "zmaxmax", "L_packet_th", "E_paquet", "T_min"):
            rec(k, [float(getattr(P, k))], 0)
        for k in GRID + OPA + EMI:
            if getattr(P, k, None) is not None:
                rec(k, getattr(P, k), 0)
        for k in GRID_I + ("l_dark_zone",):
            if getattr(P, k, None) is not None:
                rec(k, getattr(P, k), 2)
        for k in OPA_F + ("CDF_E_star",):
            if getattr(P, k, None) is not None:
                rec(k, getattr(P, k), 1)


if __name__ == "__main__":
    n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    P = S.ref41_like(n_photons_eq_th=n2, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3)
    dump(P, sys.argv[1])
    print("wrote", sys.argv[1], "n_cells", P.n_cells, "n_lambda", P.n_lambda)
