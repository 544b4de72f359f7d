"""Vectors produced by the REFERENCE's own Fortran routines (oracle/ref_harness, built by a maintainer who has gfortran)
pin the oracle when tests/golden/ref_vectors_cyl.npz is present; without the file the oracle stays "parity unpinned by the
reference" and this test is skipped."""
import os

import numpy as np
import pytest

from oracle.binding import Oracle
from helpers import small_problems

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors_cyl.npz")


@pytest.mark.skipif(not os.path.exists(PATH), reason="no reference-generated vectors (needs gfortran: make -C oracle/ref_harness vectors)")
def test_oracle_reproduces_reference_vectors():
    V = np.load(PATH)
    P = small_problems()["cyl2D"]()
    O = Oracle(P)
    o = O.cross_cell(V["x"], V["y"], V["z"], V["u"], V["v"], V["w"], V["icell"])
    assert np.array_equal(o["next_cell"], V["next_cell"])
    for k in ("x1", "y1", "z1", "l"):
        assert np.allclose(o[k], V[k], rtol=1e-12, atol=0), k
    assert np.array_equal(O.index_cell(V["x"], V["y"], V["z"]), V["index_cell"])
    assert np.allclose(O.distance_to_closest_wall(V["icell"], V["x"], V["y"], V["z"]), V["d_wall"], rtol=1e-12, atol=0)
    R = 3.0 * np.sqrt(P.Rmax2)
    m = O.move_to_grid(V["x"] - R * V["u"], V["y"] - R * V["v"], V["z"] - R * V["w"], V["u"], V["v"], V["w"])
    assert np.array_equal(m["lintersect"], V["m_lint"])
    hit = V["m_lint"] != 0
    assert np.array_equal(m["icell"][hit], V["m_icell"][hit])
    for a, b in (("x", "mx"), ("y", "my"), ("z", "mz")):
        assert np.allclose(m[a][hit], V[b][hit], rtol=1e-12, atol=0)


def test_harness_inputs_are_reproducible(tmp_path):
    """the input side of the harness runs here (the Fortran side cannot): same seeded rays every time"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_harness"))
    import make_inputs
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    make_inputs.write(a, n=500); make_inputs.write(b, n=500)
    assert open(a, "rb").read() == open(b, "rb").read() and os.path.getsize(a) > 500 * 52
