"""shim/c_driver.c: a plain-C program that fills the ABI structs field by field (as the Fortran shim does) and runs one
thermal mc_photon_loop call through libmcfost_b200.so.  Without a GPU it must fail loudly (exit 77, no fallback)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    from mcfost_b200 import build as mcb_build
    mcb_build.build()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dump_problem
    from mcfost_b200 import synthetic as S
    P = S.ref41_like(n_photons_eq_th=100, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3)
    prob = str(tmp_path / "problem.bin")
    dump_problem.dump(P, prob)
    exe = str(tmp_path / "c_driver")
    libdir = os.path.join(ROOT, "mcfost_b200", "_lib")
    subprocess.run(["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(ROOT, "shim", "c_driver.c"), "-L", libdir, "-lmcfost_b200", "-Wl,-rpath," + libdir], check=True)
    return exe, prob


def test_c_driver_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe, prob = _build(tmp_path)
    r = subprocess.run([exe, prob, "100"], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 77 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_runs_the_thermal_step(tmp_path):
    exe, prob = _build(tmp_path)
    r = subprocess.run([exe, prob, "100"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "packets 12800 " in r.stdout and "Tdust max" in r.stdout
    # the neighbours of the path, bound from C the same way: against the host mirror on the same problem
    import re
    import numpy as np
    from mcfost_b200 import synthetic as S, api
    P = S.ref41_like(n_photons_eq_th=100, dark_zone=False, n_rad=40, nz=20, n_rad_in=5, tau_mid=1.0e3)
    G = api.PhotonLoop(P)
    d = G.define_dark_zone(P.lambda_seuil, 30.0, P.r_grid, P.z_grid, [(1, P.n_rad)])
    col = G.compute_column(P.lambda_seuil, P.r_grid, np.zeros(P.n_cells), P.z_grid)
    G.close()
    m = re.search(r"dark zone: (\d+) cells, ri_in (\d+) ri_out (\d+) l_is_dark_zone (\d)", r.stdout)
    assert m and int(m.group(1)) == d["l_dark_zone"].sum() > 0 and int(m.group(2)) == d["ri_in"][0] and int(m.group(3)) == d["ri_out"][0]
    m = re.search(r"column: max optical depth ([0-9.e+-]+)", r.stdout)
    assert m and abs(float(m.group(1)) / col.max() - 1) < 1e-4
    assert "init_reemission: max |log Qcool - uploaded|" in r.stdout
