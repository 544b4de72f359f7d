"""Generate the G5 mesh (Voronoi cells of a synthetic 1M-particle SPH disk) once and cache it under
data_cache/ (git-ignored, but it travels to the GPU box with the repository snapshot).
usage: make_g5_mesh.py [n_points]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcfost_b200 import synthetic as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data_cache", "g5_%d.npz" % n)
t0 = time.time()
P = S.voronoi_sph_disk(n_points=n, cache=path)
print("G5 mesh: %d cells, %.2f neighbours per cell, %.1f %% cut cells, %.0f s -> %s" % (P.n_cells, len(P.neighbours_list) / P.n_cells, 100 * P.vor_was_cut.mean(), time.time() - t0, path))
