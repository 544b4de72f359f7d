"""ref_harness output stream -> tests/golden/ref_vectors_cyl.npz (consumed by tests/test_ref_vectors.py)."""
import sys
import numpy as np

fin, fout, npz = sys.argv[1:4]
with open(fin, "rb") as f:
    n_rad, nz, n_az, l3D, n = np.fromfile(f, np.int32, 5)
    skip = 3 * (n_rad + 1) + n_rad * (nz + 2) + n_rad + 3 * n_az + 2
    np.fromfile(f, np.float64, skip)
    x, y, z, u, v, w = (np.fromfile(f, np.float64, n) for _ in range(6))
    icell = np.fromfile(f, np.int32, n)
rec = np.dtype([("x1", "<f8"), ("y1", "<f8"), ("z1", "<f8"), ("l", "<f8"), ("next_cell", "<i4"), ("index_cell", "<i4"), ("d_wall", "<f8"),
                ("mx", "<f8"), ("my", "<f8"), ("mz", "<f8"), ("m_icell", "<i4"), ("m_lint", "<i4")])
with open(fout, "rb") as f:
    assert np.fromfile(f, np.int32, 1)[0] == n
    r = np.fromfile(f, rec, n)
np.savez(npz, x=x, y=y, z=z, u=u, v=v, w=w, icell=icell, **{k: r[k] for k in rec.names})
print("wrote", npz, n, "rays")
