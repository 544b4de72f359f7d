// Voronoi-mesh geometry, device side.
// Reference: cross_Voronoi_cell Voronoi.f90:839-992, distance_to_wall :1289-1317,
// distance_to_star :1321-1375, move_to_grid_Voronoi :1379-1442,
// test_exit_grid_Voronoi :1446-1459, is_in_volume :1463-1478,
// pos_em_cell_voronoi :1510-1543 (emits from the cell centre: the displacement
// line is commented out, :1539), index_cell_voronoi :1548-1572.
//
// The plane tests are fp32 exactly like the reference (`real, dimension(3) ::
// n, p, r, k`, :860).  Data path: seeds are read from a float4 fp32 copy
// (one 16-byte vector load per neighbour, the analogue of Voronoi_xyz :61);
// the per-cell header (first/last neighbour, flags) is read once per crossing.
#pragma once
#include "model.cuh"
#include "geom_rz.cuh"      // mc_div

namespace mcb {

struct HitVor { double l, l_contrib, l_void; int next; };
__device__ __forceinline__ double hit_l_contrib(const HitVor& h) { return h.l_contrib; }
__device__ __forceinline__ double hit_l_void(const HitVor& h) { return h.l_void; }

struct GeomVor {
  static constexpr bool is_vor = true;
  using CellT = int;
  using Hit = HitVor;

  static __device__ __forceinline__ bool test_exit(const DevModel&, int c, double, double, double) { return c < 0; }

  static __device__ __forceinline__ double distance_to_wall(const DevModel& m, double x, double y, double z,
                                                            double u, double v, double w, int iwall) {
    const double n0 = m.wall[iwall - 1][0], n1 = m.wall[iwall - 1][1], n2 = m.wall[iwall - 1][2], d = m.wall[iwall - 1][3];
    const double p0 = d * fabs(n0), p1 = d * fabs(n1), p2 = d * fabs(n2);
    const float den = (float)(n0 * u + n1 * v + n2 * w);
    if (fabsf(den) > FLT_MIN) return (n0 * (p0 - x) + n1 * (p1 - y) + n2 * (p2 - z)) / (double)den;
    return MCB_HUGE_REAL;
  }
  static __device__ __forceinline__ bool is_in_volume(const DevModel& m, double x, double y, double z) {
    return (x > m.wall[0][3]) && (x < m.wall[1][3]) && (y > m.wall[2][3]) && (y < m.wall[3][3]) && (z > m.wall[4][3]) && (z < m.wall[5][3]);
  }
  // index_cell_voronoi (Voronoi.f90:1548-1572): nearest seed, distances rounded to fp32, strict `<` in a loop over
  // ascending ids (ties go to the lowest id).  The reference scans all n_cells seeds (and finds entry cells with a
  // kd-tree, :1625-1645); here the seeds are binned in a uniform grid at upload and the search visits the shells of grid
  // cells around the point until no unvisited cell can hold a closer seed.  Same result as the O(n) scan (checked
  // against the oracle's brute force); a shell is only skipped when its nearest face is farther than the best
  // distance plus the fp32 rounding margin.
  static __device__ __forceinline__ void index_try(const DevModel& m, int id, double x, double y, double z, float& best, int& ic) {
    const double dx = __ldg(m.vor_xyz + 3 * (size_t)(id - 1)) - x, dy = __ldg(m.vor_xyz + 3 * (size_t)(id - 1) + 1) - y, dz = __ldg(m.vor_xyz + 3 * (size_t)(id - 1) + 2) - z;
    const float d2 = (float)(dx * dx + dy * dy + dz * dz);
    if (d2 < best || (d2 == best && id < ic)) { best = d2; ic = id; }
  }
  static __device__ int index(const DevModel& m, double x, double y, double z) {
    float best = FLT_MAX; int ic = 0;
    if (!m.vg_start) {      // no grid (tiny meshes): plain scan
      for (int i = 1; i <= m.n_cells; ++i) index_try(m, i, x, y, z, best, ic);
      return ic;
    }
    const double p[3] = {x, y, z};
    int c[3];
    for (int a = 0; a < 3; ++a) {
      const double q = floor((p[a] - m.vg_lo[a]) * m.vg_inv[a]);
      c[a] = q < 0.0 ? 0 : (q >= (double)m.vg_n[a] ? m.vg_n[a] - 1 : (int)q);
    }
    const int rmax = max(max(m.vg_n[0], m.vg_n[1]), m.vg_n[2]);
    for (int r = 0; r < rmax; ++r) {
      // shell r: grid cells with max(|di|, |dj|, |dk|) == r
      const int i0 = max(c[0] - r, 0), i1 = min(c[0] + r, m.vg_n[0] - 1);
      const int j0 = max(c[1] - r, 0), j1 = min(c[1] + r, m.vg_n[1] - 1);
      const int k0 = max(c[2] - r, 0), k1 = min(c[2] + r, m.vg_n[2] - 1);
      for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j) {
          const bool edge_jk = (abs(k - c[2]) == r) || (abs(j - c[1]) == r);
          const int step = edge_jk ? 1 : max(i1 - i0, 1);      // interior rows: only the two end cells belong to the shell
          for (int i = i0; i <= i1; i += step) {
            if (!edge_jk && abs(i - c[0]) != r) continue;
            const int g = i + m.vg_n[0] * (j + m.vg_n[1] * k);
            const int e1 = __ldg(m.vg_start + g + 1);
            for (int e = __ldg(m.vg_start + g); e < e1; ++e) index_try(m, __ldg(m.vg_items + e), x, y, z, best, ic);
          }
        }
      // can a cell outside the visited block [c - r, c + r] hold a closer seed?  distance from the point to the block's faces
      double dmin = 1.0e300; bool open = false;
      for (int a = 0; a < 3; ++a) {
        if (c[a] - r > 0) { open = true; dmin = fmin(dmin, p[a] - (m.vg_lo[a] + (c[a] - r) * m.vg_step[a])); }
        if (c[a] + r < m.vg_n[a] - 1) { open = true; dmin = fmin(dmin, (m.vg_lo[a] + (c[a] + r + 1) * m.vg_step[a]) - p[a]); }
      }
      if (!open) break;
      if (ic > 0 && dmin > 0.0 && dmin * dmin > (double)best * (1.0 + 1.0e-5)) break;
    }
    return ic;
  }
  static __device__ __forceinline__ double distance_to_star(const DevModel& m, double x, double y, double z, double u, double v, double w, int& i_star) {
    double dmin = MCB_HUGE_DP;
    i_star = 0;
    for (int i = 0; i < m.n_stars; ++i) {
      const double dx = x - m.star[i][0], dy = y - m.star[i][1], dz = z - m.star[i][2];
      const double b = dx * u + dy * v + dz * w;
      const double c = dx * dx + dy * dy + dz * dz - m.star[i][3] * m.star[i][3];
      const double delta = b * b - c;
      if (delta >= 0.) {
        const double rac = sqrt(delta), s1 = -b - rac;
        if (s1 < 0) { const double s2 = -b + rac; if (s2 > 0) { dmin = 0.0; i_star = i + 1; } }
        else if (s1 < dmin) { dmin = s1; i_star = i + 1; }
      }
    }
    return dmin;
  }

  // cross_Voronoi_cell (:839-992): the neighbour loop already yields the next cell
  static __device__ HitVor distance(const DevModel& m, DirInv, double x, double y, double z, double u, double v, double w,
                                    int icell, int previous_cell) {
    int next_cell; double s_contrib, s_void_before;
    const double prec = (double)1e-5f;
    const float rx = (float)x, ry = (float)y, rz = (float)z, kx = (float)u, ky = (float)v, kz = (float)w;
    double s = (double)1e30f;
    next_cell = 0;
    const float4 rc = __ldg(m.vor_xyz32 + (icell - 1));
    const int ifirst = __ldg(m.vor_first + icell - 1), ilast = __ldg(m.vor_last + icell - 1);
    const unsigned flags = __ldg(m.vor_flags + icell - 1);
    // Software pipeline, two neighbours deep: the id and the seed of the neighbours i + 1 and i + 2 are requested while
    // neighbour i is worked on (the loop is a chain of two dependent loads per neighbour -- list entry, then the seed it
    // points at -- and `long_scoreboard` was 6.3 of the 14.9 cycles per issue of this kernel on the 1M-cell mesh).  Same
    // operations on the same operands in the same order: the result is unchanged.
    int id_a = (ifirst <= ilast) ? __ldg(m.neigh + ifirst - 1) : 0;
    int id_b = (ifirst + 1 <= ilast) ? __ldg(m.neigh + ifirst) : 0;
    float4 rn_a = (id_a > 0) ? __ldg(m.vor_xyz32 + (id_a - 1)) : rc;
    for (int i = ifirst; i <= ilast; ++i) {
      const int id_n = id_a;
      const float4 rn = rn_a;
      // next iteration's operands
      id_a = id_b;
      id_b = (i + 2 <= ilast) ? __ldg(m.neigh + i + 1) : 0;
      rn_a = (id_a > 0) ? __ldg(m.vor_xyz32 + (id_a - 1)) : rc;
      if (id_n == previous_cell) continue;
      double s_tmp;
      if (id_n > 0) {
        const float nx = __fsub_rn(rn.x, rc.x), ny = __fsub_rn(rn.y, rc.y), nz = __fsub_rn(rn.z, rc.z);
        const float denf = __fadd_rn(__fadd_rn(__fmul_rn(nx, kx), __fmul_rn(ny, ky)), __fmul_rn(nz, kz));
        if (!(denf > 0.f)) continue;
        const float px = __fmul_rn(0.5f, __fadd_rn(rn.x, rc.x)), py = __fmul_rn(0.5f, __fadd_rn(rn.y, rc.y)), pz = __fmul_rn(0.5f, __fadd_rn(rn.z, rc.z));
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, __fsub_rn(px, rx)), __fmul_rn(ny, __fsub_rn(py, ry))), __fmul_rn(nz, __fsub_rn(pz, rz)));
        s_tmp = mc_div((double)dot, (double)denf);      // (IEEE division in the deterministic kernels, reciprocal + Newton in the Monte Carlo ones)
        if (s_tmp < 0.) s_tmp = MCB_HUGE_REAL;
      } else {
        s_tmp = distance_to_wall(m, x, y, z, u, v, w, -id_n);
        if (s_tmp < 0.) s_tmp = MCB_HUGE_REAL;
      }
      if (s_tmp < s) { s = s_tmp; next_cell = id_n; }
    }
    s = s * (1.0 + prec);
    if (next_cell == 0) {
      s = 0.0;
      if (is_in_volume(m, x, y, z)) {
        next_cell = index(m, x, y, z);
        if (icell == next_cell) next_cell = -1;
      } else next_cell = -1;
    }
    if (flags & 1u) {       // was_cut
      const double dx = (double)__fsub_rn(rx, rc.x), dy = (double)__fsub_rn(ry, rc.y), dz = (double)__fsub_rn(rz, rc.z);
      const double b = dx * (double)kx + dy * (double)ky + dz * (double)kz;
      const double hc = __ldg(m.vor_h + icell - 1) * m.cut_o_h;
      const double c = dx * dx + dy * dy + dz * dz - hc * hc;
      const double delta = b * b - c;
      if (delta < 0.) { s_void_before = s; s_contrib = 0.0; }
      else {
        const double rac = sqrt(delta), s1 = -b - rac, s2 = -b + rac;
        if (s1 < 0) {
          if (s2 < 0) { s_void_before = s; s_contrib = 0.0; }
          else { s_void_before = 0.0; s_contrib = fmin(s2, s); }
        } else {
          if (s1 < s) { s_void_before = s1; s_contrib = fmin(s2, s) - s1; }
          else { s_void_before = s; s_contrib = 0.0; }
        }
      }
    } else { s_void_before = 0.0; s_contrib = s; }
    if (flags & 4u) {       // is_star_neighbour
      int i_star;
      const double d_to_star = distance_to_star(m, x, y, z, u, v, w, i_star);
      if (i_star > 0 && d_to_star < s) { s_contrib = d_to_star; next_cell = m.star_icell[i_star - 1]; }
    }
    HitVor h; h.l = s; h.l_contrib = s_contrib; h.l_void = s_void_before; h.next = next_cell;
    return h;
  }

  // distance_to_closest_wall_Voronoi (Voronoi.f90:996-1061): fp32 plane geometry as in the crossing.  The reference
  // divides dot(n, p - r) by dot(n, n) with the un-normalised n = r_neighbour - r_cell, i.e. returns the distance in
  // units of |n| (dead code there); the length dot(n, p - r) / |n| is returned here.  0 for cut cells and for cells
  // that touch a wall of the box (:1009, :1051).
  static __device__ double closest_wall(const DevModel& m, int icell, double x, double y, double z) {
    const unsigned flags = __ldg(m.vor_flags + icell - 1);
    if (flags & 1u) return 0.0;
    const float rx = (float)x, ry = (float)y, rz = (float)z;
    double s = (double)1e30f;
    const float4 rc = __ldg(m.vor_xyz32 + (icell - 1));
    const int ifirst = __ldg(m.vor_first + icell - 1), ilast = __ldg(m.vor_last + icell - 1);
    for (int i = ifirst; i <= ilast; ++i) {
      const int id_n = __ldg(m.neigh + i - 1);
      double s_tmp;
      if (id_n > 0) {
        const float4 rn = __ldg(m.vor_xyz32 + (id_n - 1));
        const float nx = __fsub_rn(rn.x, rc.x), ny = __fsub_rn(rn.y, rc.y), nz = __fsub_rn(rn.z, rc.z);
        const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
        const double den = (double)__fsqrt_rn(n2);
        const float px = __fmul_rn(0.5f, __fadd_rn(rn.x, rc.x)), py = __fmul_rn(0.5f, __fadd_rn(rn.y, rc.y)), pz = __fmul_rn(0.5f, __fadd_rn(rn.z, rc.z));
        const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, __fsub_rn(px, rx)), __fmul_rn(ny, __fsub_rn(py, ry))), __fmul_rn(nz, __fsub_rn(pz, rz)));
        s_tmp = (double)dot / den;
        if (s_tmp < 0.) s_tmp = MCB_HUGE_REAL;
      } else s_tmp = 0.0;
      if (s_tmp < s) s = s_tmp;
    }
    return s;
  }

  static __device__ __forceinline__ void exit_point(const HitVor& h, double x, double y, double z, double u, double v, double w,
                                                    double& x1, double& y1, double& z1) {
    x1 = x + u * h.l; y1 = y + v * h.l; z1 = z + w * h.l;      // (h.l = 0 on the rounding fallback: the packet does not move)
  }
  static __device__ __forceinline__ void advance(const DevModel&, const HitVor& h, double x, double y, double z, double u, double v, double w,
                                                 int, double& x1, double& y1, double& z1, int& nxt) {
    exit_point(h, x, y, z, u, v, w, x1, y1, z1);
    nxt = h.next;
  }
  static __device__ double cross(const DevModel& m, DirInv d, double x, double y, double z, double u, double v, double w,
                                 int icell, int previous_cell, double& x1, double& y1, double& z1, int& next_cell,
                                 double& s_contrib, double& s_void_before) {
    HitVor h = distance(m, d, x, y, z, u, v, w, icell, previous_cell);
    advance(m, h, x, y, z, u, v, w, icell, x1, y1, z1, next_cell);
    s_contrib = h.l_contrib; s_void_before = h.l_void;
    return h.l;
  }

  static __device__ bool move_to_grid(const DevModel& m, double& x, double& y, double& z, double u, double v, double w, int& c) {
    const double prec = 1.e-6;
    double sw[6]; int order[6];
    for (int iw = 1; iw <= 6; ++iw) {
      const double l = distance_to_wall(m, x, y, z, u, v, w, iw);
      sw[iw - 1] = (l >= 0) ? l * (1.0 + prec) : MCB_HUGE_REAL;
      order[iw - 1] = iw;
    }
    for (int a = 1; a < 6; ++a) {      // stable insertion sort
      int o = order[a]; int b = a - 1;
      while (b >= 0 && sw[order[b] - 1] > sw[o - 1]) { order[b + 1] = order[b]; --b; }
      order[b + 1] = o;
    }
    for (int i = 0; i < 6; ++i) {
      const double l = sw[order[i] - 1];
      const double xt = x + l * u, yt = y + l * v, zt = z + l * w;
      if (is_in_volume(m, xt, yt, zt)) { x = xt; y = yt; z = zt; c = index(m, x, y, z); return true; }
    }
    c = 0;
    return false;
  }

  static __device__ void pos_em_cell(const DevModel& m, int c, float, float, float, double& x, double& y, double& z) {
    x = __ldg(m.vor_xyz + 3 * (size_t)(c - 1)); y = __ldg(m.vor_xyz + 3 * (size_t)(c - 1) + 1); z = __ldg(m.vor_xyz + 3 * (size_t)(c - 1) + 2);
  }
};

}  // namespace mcb
