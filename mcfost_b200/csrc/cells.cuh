// Small helpers shared by the deterministic sub-kernels (api.cu) and the Monte Carlo kernel
// (mc_kernel.cu): cell <-> id conversions for both grid families and the ray / star test.
#pragma once
#include "model.cuh"
#include "geom_rz.cuh"
#include "geom_vor.cuh"

namespace mcb {

// ---- cell helpers common to rz and Voronoi -------------------------------
__device__ __forceinline__ int tally_index(const DevModel& m, Cell c) { return is_real(m, c) ? real_index(m, c) : -1; }
__device__ __forceinline__ int tally_index(const DevModel& m, int c) { return (c >= 1 && c <= m.n_cells) ? c - 1 : -1; }
__device__ __forceinline__ bool same_cell(Cell a, Cell b) { return a.ri == b.ri && a.zj == b.zj && a.k == b.k; }
__device__ __forceinline__ bool same_cell(int a, int b) { return a == b; }
__device__ __forceinline__ void cell_of_id(const DevModel& m, int id, Cell& c) { c = cell_from_id(m, id); }
__device__ __forceinline__ void cell_of_id(const DevModel&, int id, int& c) { c = id; }
__device__ __forceinline__ int id_of_cell(const DevModel& m, Cell c) { return cell_id(m, c); }
__device__ __forceinline__ int id_of_cell(const DevModel&, int c) { return c; }
__device__ __forceinline__ void null_cell(Cell& c) { c.ri = -7; c.zj = 0; c.k = 0; }
__device__ __forceinline__ void null_cell(int& c) { c = 0; }

// ---- stars.f90:812-884 intersect_stars -> index of the star (0 = none) -----
__device__ __forceinline__ int intersect_stars(const DevModel& m, double x, double y, double z, double u, double v, double w) {
  double d_to_star = MCB_HUGE_DP;
  int i_star = 0;
  for (int i = 0; i < m.n_stars; ++i) {
    double dx = x - m.star[i][0], dy = y - m.star[i][1], dz = z - m.star[i][2];
    double b = dx * u + dy * v + dz * w;
    double c = (dx * dx + dy * dy + dz * dz) - m.star[i][3] * m.star[i][3];
    double delta = b * b - c;
    if (delta >= 0.) {
      double rac = sqrt(delta), s1 = -b - rac;
      if (s1 < 0) { double s2 = -b + rac; if (s2 > 0) { d_to_star = 0.0; i_star = i + 1; } }
      else if (s1 < d_to_star) { d_to_star = s1; i_star = i + 1; }
    }
  }
  return i_star;
}

}  // namespace mcb
