// Structured-grid geometry (cylindrical + spherical), device side.
//
// B200-native restructuring of the reference routines -- same floating-point
// operation order (results are bit-identical to the Fortran logic when built
// with --fmad=false), different data path:
//   * cells are carried as (ri, zj, k) triples in registers; the reference's
//     cell_map / cell_map_i/j/k gathers (cylindrical_grid.f90:34-35, 4 dependent
//     loads per step) are replaced by the closed-form numbering below, which
//     reproduces build_cylindrical_cell_mapping (cylindrical_grid.f90:45-179);
//   * direction-only terms (1/(u^2+v^2), 1/w) are hoisted out of the per-cell
//     loop (the reference's own "TODO: can be calculated outside", :937-953).
//
// Reference: cross_cylindrical_cell cylindrical_grid.f90:918-1175,
// index_cell_cyl :833-890, test_exit_grid_cyl :680-704, move_to_grid_cyl
// :1284-1411, pos_em_cell_cyl :1415-1466; cross_spherical_cell
// spherical_grid.f90:182-446, index_cell_sph :48-125, move_to_grid_sph :562-615,
// pos_em_cell_sph :619-699, test_exit_grid_sph :24-44.
#pragma once
#include "model.cuh"

namespace mcb {

struct Cell { int ri, zj, k; };

// Reciprocal / division of the Monte Carlo kernels.  mc_kernel.cu defines MCB_MC_FAST_MATH: MUFU seed (>= 20 bits) + two
// Newton steps, 5 instructions and <= 2 ulp, instead of the IEEE division sequence (~25 instructions and a slow-path
// call) -- the photon loop is compared statistically and already contracts FMAs.  api.cu (deterministic kernels, results
// bit-identical to the reference's arithmetic) does not define it: IEEE division.  Arguments are normal, non-zero numbers.
#ifdef MCB_MC_FAST_MATH
__device__ __forceinline__ double mc_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
__device__ __forceinline__ double mc_div(double a, double b) { return a * mc_rcp(b); }
#else
__device__ __forceinline__ double mc_rcp(double x) { return 1.0 / x; }
__device__ __forceinline__ double mc_div(double a, double b) { return a / b; }
#endif

__device__ __forceinline__ float max_int_f() { return (float)2147483647 * (1.0f - 1.0e-5f); }   // constants.f90:159

// Fortran MODULO(a, p) for p > 0
__device__ __forceinline__ double fmodulo(double a, double p) {
  double r = fmod(a, p);
  if (r != 0.0 && r < 0.0) r += p;
  return r;
}

// ---------------- closed-form cell numbering ------------------------------
__device__ __host__ __forceinline__ int row_index(const DevModel& m, int j) {   // 0-based among real rows
  return m.l3D ? (j < 0 ? j + m.nz : j + m.nz - 1) : j - 1;
}
__device__ __host__ __forceinline__ bool is_real(const DevModel& m, Cell c) {
  int aj = c.zj < 0 ? -c.zj : c.zj;
  return c.ri >= 1 && c.ri <= m.n_rad && aj >= 1 && aj <= m.nz;
}
// linear 0-based id of a REAL cell (tally index)
__device__ __host__ __forceinline__ int real_index(const DevModel& m, Cell c) {
  return (c.ri - 1) + m.n_rad * (row_index(m, c.zj) + m.nj * (c.k - 1));
}
// reference 1-based id (real or virtual)
__device__ __host__ __forceinline__ int cell_id(const DevModel& m, Cell c) {
  if (is_real(m, c)) return real_index(m, c) + 1;
  const int jstart2 = m.l3D ? -m.nz - 1 : 0, jend2 = m.nz + 1;
  if (c.zj == jstart2 || c.zj == jend2) {
    int row = (c.zj == jstart2) ? 0 : 1;
    return m.n_cells + (c.k - 1) * 2 * (m.n_rad + 2) + row * (m.n_rad + 2) + c.ri + 1;
  }
  int base2 = m.n_cells + m.n_az * 2 * (m.n_rad + 2);
  return base2 + (c.k - 1) * 2 * m.nj + row_index(m, c.zj) * 2 + (c.ri == 0 ? 1 : 2);
}
__device__ __host__ __forceinline__ Cell cell_from_id(const DevModel& m, int id) {
  Cell c;
  if (id <= m.n_cells) {
    int q = id - 1;
    c.ri = q % m.n_rad + 1; q /= m.n_rad;
    int jr = q % m.nj; c.k = q / m.nj + 1;
    c.zj = m.l3D ? (jr < m.nz ? jr - m.nz : jr - m.nz + 1) : jr + 1;
    return c;
  }
  const int jstart2 = m.l3D ? -m.nz - 1 : 0, jend2 = m.nz + 1;
  int base2 = m.n_cells + m.n_az * 2 * (m.n_rad + 2);
  if (id <= base2) {
    int q = id - m.n_cells - 1;
    c.ri = q % (m.n_rad + 2); q /= (m.n_rad + 2);
    c.zj = (q % 2 == 0) ? jstart2 : jend2; c.k = q / 2 + 1;
    return c;
  }
  int q = id - base2 - 1;
  c.ri = (q % 2 == 0) ? 0 : m.n_rad + 1; q /= 2;
  int jr = q % m.nj; c.k = q / m.nj + 1;
  c.zj = m.l3D ? (jr < m.nz ? jr - m.nz : jr - m.nz + 1) : jr + 1;
  return c;
}

// ---------------- table accessors (1-based like the reference) -------------
// SM = true: the table lives in the block's shared-memory staging area (SmemLayout);
// SM = false: read-only global loads.
extern __shared__ __align__(16) unsigned char mcb_smem_raw[];
__device__ __forceinline__ const double* smd() { return reinterpret_cast<const double*>(mcb_smem_raw); }
// Read-only loads from the tables staged in shared memory.  (Measured and dropped in round 2: nvcc rebuilds the address of
// the dynamic shared array from %cluster_ctarank on every access, S2R + MOV + LEA in front of the LDS; replacing that by
// `ld.shared` inline asm on a host-provided window base removed the three instructions but, being opaque to the scheduler,
// cost the packet-per-lane kernel 5 % and gained the packet-per-warp kernel 1 %.)
__device__ __forceinline__ double smd_ld(const DevModel&, int word) { return smd()[word]; }
__device__ __forceinline__ float smf_ld(const DevModel&, int word, int k) { return reinterpret_cast<const float*>(smd() + word)[k]; }
template <bool SM> __device__ __forceinline__ double r_lim_2(const DevModel& m, int i) { return SM ? smd_ld(m, m.sm.r_lim_2 + i) : __ldg(m.r_lim_2 + i); }
template <bool SM> __device__ __forceinline__ double zmax(const DevModel& m, int i) { return SM ? smd_ld(m, m.sm.zmax + i - 1) : __ldg(m.zmax + (i - 1)); }
// z_lim(i,j), j = 1..nz+1.  For the default grid z_lim(i,j) = (j-1)*cell_height(i) and z_lim(i,nz+1) = zmax(i)
// (cylindrical_grid.f90:458-465); upload_grid verifies that bit-for-bit and then only cell_height is kept.
template <bool SM> __device__ __forceinline__ double z_lim(const DevModel& m, int i, int j) {
  if (m.z_regular) {
    if (j == m.nz + 1) return zmax<SM>(m, i);
    const double ch = SM ? smd_ld(m, m.sm.zl + i - 1) : __ldg(m.cell_height + (i - 1));
    return (double)(j - 1) * ch;
  }
  return SM ? smd_ld(m, m.sm.zl + (i - 1) + m.n_rad * (j - 1)) : __ldg(m.z_lim + (i - 1) + m.n_rad * (j - 1));
}
template <bool SM> __device__ __forceinline__ double tan_phi_lim(const DevModel& m, int k) { return SM ? smd_ld(m, m.sm.tan_phi + k - 1) : __ldg(m.tan_phi_lim + (k - 1)); }
template <bool SM> __device__ __forceinline__ double tan_theta_lim(const DevModel& m, int j) { return SM ? smd_ld(m, m.sm.tan_theta + j) : __ldg(m.tan_theta_lim + j); }

// result of the wall-distance half of a crossing (the next-cell half is only needed when the flight goes on)
struct HitRZ {
  double l;        // = l_contrib (l_void_before = 0 on structured grids)
  int which;       // 0 radial, 1 vertical / theta, 2 azimuthal
  int d_rad, d_j, d_phi;
};

__device__ __forceinline__ double hit_l_contrib(const HitRZ& h) { return h.l; }
__device__ __forceinline__ double hit_l_void(const HitRZ&) { return 0.0; }

// direction-only invariants of a flight
struct DirInv { double inv_a, inv_w; };
__device__ __forceinline__ DirInv dir_invariants(double u, double v, double w) {
  DirInv d;
  double a = u * u + v * v;
  d.inv_a = (a > MCB_TINY_REAL) ? mc_rcp(a) : MCB_HUGE_REAL;
  d.inv_w = (fabs(w) > MCB_TINY_REAL) ? mc_rcp(w) : copysign(MCB_HUGE_DP, w);
  return d;
}

__device__ __forceinline__ int phi_index(const DevModel& m, double x, double y) {
  double phi = fmodulo(atan2(y, x), 2 * MCB_PI);
  int k = (int)floor(phi / (2 * MCB_PI) * (double)(float)m.n_az) + 1;
  if (k == m.n_az + 1) k = m.n_az;
  return k;
}

// sin_phi_lim(k) as a wall normal: the grid set-up stores 1.0d300 next to tan_phi_lim = 1.0d300 for the walls at
// phi = pi/2 (mod pi) (cylindrical_grid.f90:591-594), which would take them out of distance_to_closest_wall's minimum;
// they are the planes x = 0 (sin = 1, cos = 0)
__device__ __forceinline__ double sin_phi_wall(const DevModel& m, int k) { const double sp = __ldg(m.sin_phi_lim + k - 1); return sp > 1.0e299 ? 1.0 : sp; }

// =========================================================================
// cylindrical
// =========================================================================
template <bool L3D, bool SM = false>
struct GeomCyl {
  using Hit = HitRZ;
  static constexpr bool is_vor = false;
  using CellT = Cell;

  static __device__ __forceinline__ bool test_exit(const DevModel& m, Cell c, double /*x*/, double /*y*/, double z) {
    if (is_real(m, c)) return false;
    if (c.ri == m.n_rad + 1) return true;                          // lexit_cell == 1
    int aj = c.zj < 0 ? -c.zj : c.zj;
    if (aj == m.nz + 1) return fabs(z) > m.zmaxmax;                // lexit_cell == 2
    return false;
  }

  static __device__ __forceinline__ int z_index_f32(const DevModel& m, double z, int ri) {
    // floor(min(real(abs(z)/zmax(ri)*nz), max_int)) + 1   (cylindrical_grid.f90:868,1116: fp32 cast)
    float q = (float)(mc_div(fabs(z), zmax<SM>(m, ri)) * m.nz);
    return (int)floorf(fminf(q, max_int_f())) + 1;
  }

  static __device__ Cell index(const DevModel& m, double x, double y, double z) {
    Cell c;
    double r2 = x * x + y * y;
    if (r2 < r_lim_2<SM>(m, 0)) { c.ri = 0; c.zj = 1; c.k = 1; return c; }
    if (r2 > m.Rmax2) { c.ri = m.n_rad + 1; c.zj = 1; c.k = 1; return c; }
    int lo = 0, hi = m.n_rad, ri = (lo + hi) / 2;
    while (hi - lo > 1) {
      if (r2 > r_lim_2<SM>(m, ri)) lo = ri; else hi = ri;
      ri = (lo + hi) / 2;
    }
    c.ri = ri + 1;
    int zj = z_index_f32(m, z, c.ri);
    if (zj > m.nz) zj = m.nz + 1;
    c.k = 1;
    if (L3D) {
      if (z < 0.0) zj = -zj;
      if (z != 0.0) c.k = phi_index(m, x, y);
    }
    c.zj = zj;
    return c;
  }

#ifdef MCB_BRANCHY_DISTANCE      // round-1 form, kept for A/B timing
  // ---- wall-distance half of cross_cylindrical_cell (:941-1094): which wall, how far ----
  static __device__ __forceinline__ HitRZ distance(const DevModel& m, DirInv d, double x0, double y0, double z0,
                                                   double u, double v, double w, Cell c, Cell /*prev*/) {
    const double correct_moins = 1.0 - MCB_GRID_PREC, correct_plus = 1.0 + MCB_GRID_PREC;
    const int ri0 = c.ri, zj0 = c.zj, k0 = c.k;
    double s, t, t_phi;
    HitRZ h; h.d_rad = 1; h.d_j = 0; h.d_phi = 0;
    const double r_2 = x0 * x0 + y0 * y0;
    const double b = (x0 * u + y0 * v) * d.inv_a;

    if (ri0 == 0) {
      double cc = (r_2 - r_lim_2<SM>(m, 0)) * d.inv_a;
      double rac = sqrt(b * b - cc);
      s = (-b + rac) * correct_plus;
      t = MCB_HUGE_REAL; t_phi = MCB_HUGE_REAL;
    } else {
      // 1) radial wall
      double dotprod = u * x0 + v * y0, delta;
      if (dotprod < 0.0) {
        double cc = (r_2 - r_lim_2<SM>(m, ri0 - 1) * correct_moins) * d.inv_a;
        delta = b * b - cc;
        if (delta < 0.0) {
          cc = (r_2 - r_lim_2<SM>(m, ri0) * correct_plus) * d.inv_a;
          delta = fmax(b * b - cc, 0.0);
        } else h.d_rad = -1;
      } else {
        double cc = (r_2 - r_lim_2<SM>(m, ri0) * correct_plus) * d.inv_a;
        delta = fmax(b * b - cc, 0.0);
      }
      double rac = sqrt(delta);
      s = (-b - rac) * correct_plus;
      if (s < 0.0) s = (-b + rac) * correct_plus;
      else if (s == 0.0) s = MCB_GRID_PREC;

      // 2) horizontal wall
      dotprod = w * z0;
      if (dotprod == 0.0) t = (double)1.0e10f;
      else {
        const int aj = zj0 < 0 ? -zj0 : zj0;
        double zlim;
        if (dotprod > 0.0) {
          if (aj == m.nz + 1) { h.d_j = 0; zlim = copysign(1.0e10, z0); }
          else {
            zlim = copysign(z_lim<SM>(m, ri0, aj + 1) * correct_plus, z0);
            h.d_j = (L3D && z0 < 0.0) ? -1 : 1;
          }
        } else {
          if (L3D) {
            if (z0 > 0.0) { zlim = z_lim<SM>(m, ri0, aj) * correct_moins; h.d_j = (zj0 == 1) ? -2 : -1; }
            else { zlim = -z_lim<SM>(m, ri0, aj) * correct_moins; h.d_j = (zj0 == -1) ? 2 : 1; }
          } else {
            if (zj0 == 1) {          // midplane mirror in 2D
              h.d_j = 1;
              zlim = (z0 > 0.0) ? -z_lim<SM>(m, ri0, 2) * correct_moins : z_lim<SM>(m, ri0, 2) * correct_moins;
            } else {
              zlim = (z0 > 0.0) ? z_lim<SM>(m, ri0, zj0) * correct_moins : -z_lim<SM>(m, ri0, zj0) * correct_moins;
              h.d_j = -1;
            }
          }
        }
        t = (zlim - z0) * d.inv_w;
        if (t < 0.0) t = MCB_GRID_PREC;
      }

      // 3) azimuthal wall
      if (L3D) {
        dotprod = x0 * v - y0 * u;
        if (fabs(dotprod) < (double)1.0e-10f) t_phi = (double)1.0e30f;
        else {
          double tan_angle_lim;
          if (dotprod > 0.0) { tan_angle_lim = tan_phi_lim<SM>(m, k0); h.d_phi = 1; }
          else { int km = k0 - 1; if (km == 0) km = m.n_az; tan_angle_lim = tan_phi_lim<SM>(m, km); h.d_phi = -1; }
          if (tan_angle_lim > 1.0e299) t_phi = (fabs(u) > (double)1e-6f) ? -x0 / u : (double)1.0e30f;
          else {
            double den = v - u * tan_angle_lim;
            t_phi = (fabs(den) > (double)1.0e-6f) ? -(y0 - x0 * tan_angle_lim) / den : (double)1.0e30f;
          }
          if (t_phi < 0.0) t_phi = (double)1.0e30f;
        }
      } else t_phi = MCB_HUGE_REAL;
    }
    if ((s < t) && (s < t_phi)) { h.l = s; h.which = 0; }
    else if (t < t_phi) { h.l = t; h.which = 1; }
    else { h.l = t_phi; h.which = 2; }
    return h;
  }

#else
  // ---- wall-distance half of cross_cylindrical_cell (:941-1094): which wall, how far ----
  // Written with selects instead of the reference's nested ifs: a warp's 32 packets sit in different cells and fly in
  // different directions, so every two-sided test of the Fortran (inward / outward, up / down, above / below the
  // midplane) diverges.  Each candidate is formed with the reference's own operations in the reference's order and the
  // result is chosen afterwards, so the values are bit-identical to the branched form (checked against the oracle).
  static __device__ __forceinline__ HitRZ distance(const DevModel& m, DirInv d, double x0, double y0, double z0,
                                                   double u, double v, double w, Cell c, Cell /*prev*/) {
    const double correct_moins = 1.0 - MCB_GRID_PREC, correct_plus = 1.0 + MCB_GRID_PREC;
    const int ri0 = c.ri, zj0 = c.zj, k0 = c.k;
    double s, t, t_phi;
    HitRZ h; h.d_rad = 1; h.d_j = 0; h.d_phi = 0;
    const double r_2 = x0 * x0 + y0 * y0;
    const double b = (x0 * u + y0 * v) * d.inv_a;

    if (ri0 == 0) {
      double cc = (r_2 - r_lim_2<SM>(m, 0)) * d.inv_a;
      double rac = sqrt(b * b - cc);
      s = (-b + rac) * correct_plus;
      t = MCB_HUGE_REAL; t_phi = MCB_HUGE_REAL;
    } else {
      // 1) radial wall: inner wall if the packet moves inwards and its line reaches it, else the outer wall
      {
        const double dotprod = u * x0 + v * y0;
        const double bb = b * b;
        const double delta_in = bb - (r_2 - r_lim_2<SM>(m, ri0 - 1) * correct_moins) * d.inv_a;
        const double delta_out = fmax(bb - (r_2 - r_lim_2<SM>(m, ri0) * correct_plus) * d.inv_a, 0.0);
        const bool inner = (dotprod < 0.0) && !(delta_in < 0.0);
        h.d_rad = inner ? -1 : 1;
        const double rac = sqrt(inner ? delta_in : delta_out);
        const double sm = (-b - rac) * correct_plus, sp = (-b + rac) * correct_plus;
        s = (sm < 0.0) ? sp : ((sm == 0.0) ? MCB_GRID_PREC : sm);
      }
      // 2) horizontal wall
      {
        const double dotprod = w * z0;
        const int aj = zj0 < 0 ? -zj0 : zj0;
        const bool up = dotprod > 0.0;                     // away from the midplane
        const bool top = aj == m.nz + 1;
        const bool zpos = z0 > 0.0;
        // wall row and the side of the midplane it is taken on
        int jw, dj; bool neg;
        if (L3D) {
          jw = up ? aj + 1 : aj;
          neg = up ? signbit(z0) : !zpos;
          dj = up ? ((z0 < 0.0) ? -1 : 1) : (zpos ? ((zj0 == 1) ? -2 : -1) : ((zj0 == -1) ? 2 : 1));
        } else {
          const bool mirror = zj0 == 1;                    // 2D: the midplane reflects (:1031-1050)
          jw = up ? aj + 1 : (mirror ? 2 : zj0);
          neg = up ? signbit(z0) : (mirror ? zpos : !zpos);
          dj = up ? 1 : (mirror ? 1 : -1);
        }
        if (up && top) { dj = 0; jw = aj; }
        const double zl = z_lim<SM>(m, ri0, jw) * (up ? correct_plus : correct_moins);
        double zlim = neg ? -zl : zl;
        if (up && top) zlim = copysign(1.0e10, z0);
        t = (zlim - z0) * d.inv_w;
        if (t < 0.0) t = MCB_GRID_PREC;
        if (dotprod == 0.0) { t = (double)1.0e10f; dj = 0; }
        h.d_j = dj;
      }
      // 3) azimuthal wall
      if (L3D) {
        const double dotprod = x0 * v - y0 * u;
        const bool fwd = dotprod > 0.0;
        int kw = fwd ? k0 : k0 - 1; if (kw == 0) kw = m.n_az;
        const bool par = fabs(dotprod) < (double)1.0e-10f;
        h.d_phi = par ? 0 : (fwd ? 1 : -1);
        const double tan_angle_lim = tan_phi_lim<SM>(m, kw);
        const double den = v - u * tan_angle_lim;
        const bool plane_x0 = tan_angle_lim > 1.0e299;          // wall at phi = pi/2 (mod pi): the plane x = 0
        const double num = plane_x0 ? -x0 : -(y0 - x0 * tan_angle_lim), dsel = plane_x0 ? u : den;
        t_phi = (fabs(dsel) > (double)1.0e-6f) ? mc_div(num, dsel) : (double)1.0e30f;      // (one division for both forms)
        if (t_phi < 0.0) t_phi = (double)1.0e30f;
        if (par) t_phi = (double)1.0e30f;
      } else t_phi = MCB_HUGE_REAL;
    }
    if ((s < t) && (s < t_phi)) { h.l = s; h.which = 0; }
    else if (t < t_phi) { h.l = t; h.which = 1; }
    else { h.l = t_phi; h.which = 2; }
    return h;
  }

#endif
  // exit point on the wall (:1100-1165)
  static __device__ __forceinline__ void exit_point(const HitRZ& h, double x0, double y0, double z0, double u, double v, double w,
                                                    double& x1, double& y1, double& z1) {
    const double dv = (h.which == 2) ? (1.0 + MCB_GRID_PREC) * h.l : h.l;
    x1 = x0 + dv * u; y1 = y0 + dv * v; z1 = z0 + dv * w;
  }

  // ---- next-cell half (:1106-1168): only needed when the flight continues past the wall ----
  static __device__ __forceinline__ void advance(const DevModel& m, const HitRZ& h, double x0, double y0, double z0,
                                                 double u, double v, double w, Cell c,
                                                 double& x1, double& y1, double& z1, Cell& nxt) {
    const int ri0 = c.ri, zj0 = c.zj, k0 = c.k;
    exit_point(h, x0, y0, z0, u, v, w, x1, y1, z1);
    if (h.which == 0) {
      nxt.ri = ri0 + h.d_rad;
      if (nxt.ri == 0) { nxt.zj = 1; nxt.k = 1; }
      else {
        if (nxt.ri > m.n_rad) nxt.zj = zj0;
        else {
          int zj1 = z_index_f32(m, z1, nxt.ri);
          if (zj1 > m.nz) zj1 = m.nz + 1;
          if (L3D && (z1 < 0.0)) zj1 = -zj1;
          nxt.zj = zj1;
        }
        nxt.k = k0;
        if (L3D && ri0 == 0) {
          double phi = fmodulo(atan2(y1, x1), 2 * MCB_PI);
          int k1 = (int)floor(phi * (1.0 / MCB_TWO_PI) * (double)(float)m.n_az) + 1;
          if (k1 == m.n_az + 1) k1 = m.n_az;
          nxt.k = k1;
        }
      }
    } else if (h.which == 1) {
      nxt.ri = ri0; nxt.zj = zj0 + h.d_j; nxt.k = k0;
    } else {
      nxt.ri = ri0;
      int zj1 = (int)floor(mc_div(fabs(z1), zmax<SM>(m, ri0)) * m.nz) + 1;       // fp64 here (:1150)
      if (zj1 > m.nz) zj1 = m.nz + 1;
      if (z1 < 0.0) zj1 = -zj1;
      nxt.zj = zj1;
      int k1 = k0 + h.d_phi;
      if (k1 == 0) k1 = m.n_az;
      if (k1 == m.n_az + 1) k1 = 1;
      nxt.k = k1;
    }
    if (z1 == 0.0) z1 = L3D ? copysign(MCB_GRID_PREC, w) : MCB_GRID_PREC;
  }

  // distance_to_closest_wall_cyl (cylindrical_grid.f90:1179-1226).  The wall below azimuthal sector 1 is the upper
  // wall of sector n_az (the reference indexes sin_phi_lim(0) there, out of bounds).
  static __device__ double closest_wall(const DevModel& m, Cell c, double x, double y, double z) {
    const int ri0 = c.ri, aj = c.zj < 0 ? -c.zj : c.zj, k0 = c.k;
    const double r = sqrt(x * x + y * y);
    const double s1 = __ldg(m.r_lim + ri0) - r;
    const double s2 = r - __ldg(m.r_lim + ri0 - 1);
    const double z0 = fabs(z);
    const double s3 = z_lim<SM>(m, ri0, aj + 1) - z0;
    const double s4 = z0 - z_lim<SM>(m, ri0, aj);
    double s = fmin(fmin(s1, s2), fmin(s3, s4));
    if (L3D) {
      const int km = (k0 - 1 >= 1) ? k0 - 1 : m.n_az;
      const double s5 = fabs(x * sin_phi_wall(m, k0) - y * __ldg(m.cos_phi_lim + k0 - 1));
      const double s6 = fabs(x * sin_phi_wall(m, km) - y * __ldg(m.cos_phi_lim + km - 1));
      s = fmin(s, fmin(s5, s6));
    }
    return s;
  }

  // the full cross_cylindrical_cell (deterministic kernels)
  static __device__ __forceinline__ double cross(const DevModel& m, DirInv d, double x0, double y0, double z0,
                                                 double u, double v, double w, Cell c, Cell prev,
                                                 double& x1, double& y1, double& z1, Cell& nxt,
                                                 double& l_contrib, double& l_void) {
    HitRZ h = distance(m, d, x0, y0, z0, u, v, w, c, prev);
    advance(m, h, x0, y0, z0, u, v, w, c, x1, y1, z1, nxt);
    l_contrib = h.l; l_void = 0.0;
    return h.l;
  }

  static __device__ bool move_to_grid(const DevModel& m, double& x, double& y, double& z, double u, double v, double w, Cell& c) {
    const double correct_moins = 1.0 - 1.0e-10;
    double x0 = x, y0 = y, z0 = z;
    DirInv d = dir_invariants(u, v, w);
    double r_2 = x0 * x0 + y0 * y0;
    double b = (x0 * u + y0 * v) * d.inv_a;
    double cc = (r_2 - r_lim_2<SM>(m, m.n_rad) * correct_moins) * d.inv_a;
    double delta = b * b - cc, s1, s2, t1, t2, delta_vol;
    if (delta < 0.0) { s1 = MCB_HUGE_REAL; s2 = MCB_HUGE_REAL; }
    else { double rac = sqrt(delta); s1 = -b - rac; s2 = -b + rac; }
    double dotprod = w * z0;
    if (fabs(dotprod) < MCB_TINY_REAL) { t1 = MCB_HUGE_REAL; t2 = MCB_HUGE_REAL; }
    else {
      double zlim = m.zmaxmax * correct_moins, zlim2 = -(m.zmaxmax * correct_moins);
      if (!(z0 > 0.0)) { double tmp = zlim; zlim = zlim2; zlim2 = tmp; }
      t1 = (zlim - z0) * d.inv_w; t2 = (zlim2 - z0) * d.inv_w;
    }
    if (t1 > (double)1e20f && s1 > (double)1e20f) return false;
    if (t1 > s1) {
      if (t1 > s2) {
        delta_vol = s1;
        double z1 = z0 + delta_vol * w;
        if (fabs(z1) > m.zmaxmax) return false;
      } else delta_vol = t1;
    } else {
      if (t2 < s1) return false;
      delta_vol = s1;
    }
    x = x0 + delta_vol * u; y = y0 + delta_vol * v; z = z0 + delta_vol * w;
    c = index(m, x, y, z);
    return true;
  }

  static __device__ void pos_em_cell(const DevModel& m, Cell c, float rand1, float rand2, float rand3, double& x, double& y, double& z) {
    const int ri = c.ri, zj = c.zj;
    double r = sqrt(r_lim_2<SM>(m, ri - 1) + rand1 * (r_lim_2<SM>(m, ri) - r_lim_2<SM>(m, ri - 1)));
    if (L3D) {
      if (zj > 0) z = z_lim<SM>(m, ri, zj) + rand2 * (z_lim<SM>(m, ri, zj + 1) - z_lim<SM>(m, ri, zj));
      else z = -(z_lim<SM>(m, ri, -zj) + rand2 * (z_lim<SM>(m, ri, -zj + 1) - z_lim<SM>(m, ri, -zj)));
    } else {
      if (rand2 > 0.5) z = z_lim<SM>(m, ri, zj) + (2.0 * (rand2 - 0.5)) * (z_lim<SM>(m, ri, zj + 1) - z_lim<SM>(m, ri, zj));
      else z = -(z_lim<SM>(m, ri, zj) + (2.0 * rand2) * (z_lim<SM>(m, ri, zj + 1) - z_lim<SM>(m, ri, zj)));
    }
    double phi = 2.0 * MCB_PI * ((double)c.k - 1.0 + rand3) / (double)m.n_az;
    double sp, cp; sincos(phi, &sp, &cp);
    x = r * cp; y = r * sp;
  }
};

// =========================================================================
// spherical
// =========================================================================
template <bool L3D, bool SM = false>
struct GeomSph {
  using Hit = HitRZ;
  static constexpr bool is_vor = false;
  using CellT = Cell;

  static __device__ __forceinline__ bool test_exit(const DevModel& m, Cell c, double, double, double) {
    return (!is_real(m, c)) && (c.ri == m.n_rad + 1);
  }

  static __device__ void theta_phi_index(const DevModel& m, double x, double y, double z, int& tj, int& pk) {
    double r02 = x * x + y * y;
    double tan_theta = (r02 > MCB_TINY_DP) ? fabs(z) / sqrt(r02) : (double)1.0e30f;
    int lo = 0, hi = m.nz, j = (lo + hi) / 2;
    while (hi - lo > 1) {
      if (tan_theta > tan_theta_lim<SM>(m, j)) lo = j; else hi = j;
      j = (lo + hi) / 2;
    }
    tj = j + 1; pk = 1;
    if (L3D) {
      if (z < 0) tj = -tj;
      if (z != 0.0) pk = phi_index(m, x, y);
    }
  }

  static __device__ Cell index(const DevModel& m, double x, double y, double z) {
    Cell c;
    double r2 = x * x + y * y + z * z;
    // note: the reference forms r2 = (x*x+y*y) + z*z
    r2 = (x * x + y * y) + z * z;
    if (r2 < r_lim_2<SM>(m, 0)) { c.ri = 0; c.zj = 1; c.k = 1; return c; }
    if (r2 > m.Rmax2) { c.ri = m.n_rad + 1; c.zj = 1; c.k = 1; return c; }
    int lo = 0, hi = m.n_rad, ri = (lo + hi) / 2;
    while (hi - lo > 1) {
      if (r2 > r_lim_2<SM>(m, ri)) lo = ri; else hi = ri;
      ri = (lo + hi) / 2;
    }
    c.ri = ri + 1;
    theta_phi_index(m, x, y, z, c.zj, c.k);
    return c;
  }

  static __device__ __forceinline__ double cone_root(double tan_lim, double x0, double y0, double z0, double u, double v, double w) {
    const double precision = 1.0e-15;
    double tan2 = tan_lim * tan_lim;
    double a_theta = w * w - tan2 * (u * u + v * v);
    double a_theta_m1 = 1.0 / a_theta;
    double b_theta = w * z0 - tan2 * (x0 * u + y0 * v);
    double c_theta = z0 * z0 - tan2 * (x0 * x0 + y0 * y0);
    double delta = b_theta * b_theta - a_theta * c_theta;
    if (delta < 0.0) return 1.0e30;
    double rac = sqrt(delta);
    double r1 = (-b_theta - rac) * a_theta_m1, r2 = (-b_theta + rac) * a_theta_m1;
    if (r1 <= precision) return (r2 <= precision) ? 1.0e30 : r2;
    return (r2 <= precision) ? r1 : fmin(r1, r2);
  }

  // ---- wall-distance half of cross_spherical_cell (:205-381) ----
  static __device__ __forceinline__ HitRZ distance(const DevModel& m, DirInv /*d*/, double x0, double y0, double z0,
                                                   double u, double v, double w, Cell c, Cell /*prev*/) {
    const double correct_moins = 1.0 - MCB_PREC_SPH, correct_plus = 1.0 + MCB_PREC_SPH;
    const int ri0 = c.ri, tj0 = c.zj, pk0 = c.k;
    const int atj = tj0 < 0 ? -tj0 : tj0;
    double s, t, t_phi;
    HitRZ h; h.d_rad = 1; h.d_j = 0; h.d_phi = 0;
    const double r0_2 = (x0 * x0 + y0 * y0) + z0 * z0;
    const double b = (x0 * u + y0 * v + z0 * w);
    if (ri0 == 0) {
      double cc = (r0_2 - r_lim_2<SM>(m, 0) * correct_plus);
      double rac = sqrt(b * b - cc);
      s = (-b + rac) * correct_plus;
      t = MCB_HUGE_REAL; t_phi = MCB_HUGE_REAL;
    } else {
      double delta;
      if (b < 0.0) {
        double cc = (r0_2 - r_lim_2<SM>(m, ri0 - 1) * correct_moins);
        delta = b * b - cc;
        if (delta < 0.0) { cc = (r0_2 - r_lim_2<SM>(m, ri0) * correct_plus); delta = fmax(b * b - cc, 0.0); }
        else h.d_rad = -1;
      } else {
        double cc = (r0_2 - r_lim_2<SM>(m, ri0) * correct_plus);
        delta = fmax(b * b - cc, 0.0);
      }
      double rac = sqrt(delta);
      s = -b - rac;
      if (s < 0.0) s = -b + rac; else if (s == 0.0) s = MCB_GRID_PREC;

      double lim1 = tan_theta_lim<SM>(m, atj) * correct_plus, lim2 = tan_theta_lim<SM>(m, atj - 1) * correct_moins;
      if (!(z0 >= 0.0)) { lim1 = -lim1; lim2 = -lim2; }
      double t1 = cone_root(lim1, x0, y0, z0, u, v, w);
      double t2 = cone_root(lim2, x0, y0, z0, u, v, w);
      if (t1 < t2) { t = t1; h.d_j = (atj == m.nz) ? 0 : 1; }
      else { t = t2; h.d_j = (atj == 1) ? 0 : -1; }

      if (L3D) {
        double dotprod = x0 * v - y0 * u;
        if (fabs(dotprod) < (double)1.0e-10f) { t_phi = (double)1.0e30f; h.d_phi = 0; }
        else {
          double tan_angle_lim;
          if (dotprod > 0.0) { tan_angle_lim = tan_phi_lim<SM>(m, pk0); h.d_phi = 1; }
          else { int km = pk0 - 1; if (km == 0) km = m.n_az; tan_angle_lim = tan_phi_lim<SM>(m, km); h.d_phi = -1; }
          if (tan_angle_lim > 1.0e299) t_phi = -x0 / u;
          else {
            double den = v - u * tan_angle_lim;
            if (fabs(den) > (double)1.0e-6f) t_phi = -(y0 - x0 * tan_angle_lim) / den;
            else { t_phi = (double)1.0e30f; h.d_phi = 0; }
          }
          if (t_phi < 0.0) { t_phi = (double)1.0e30f; h.d_phi = 0; }
        }
      } else t_phi = MCB_HUGE_REAL;
    }
    if ((s < t) && (s < t_phi)) { h.l = s; h.which = 0; }
    else if (t < t_phi) { h.l = t; h.which = 1; }
    else { h.l = t_phi; h.which = 2; }
    return h;
  }

  static __device__ __forceinline__ void exit_point(const HitRZ& h, double x0, double y0, double z0, double u, double v, double w,
                                                    double& x1, double& y1, double& z1) {
    const double dv = (h.which == 2) ? (1.0 + MCB_PREC_SPH) * h.l : h.l;
    x1 = x0 + dv * u; y1 = y0 + dv * v; z1 = z0 + dv * w;
  }

  // ---- next-cell half (:384-439) ----
  static __device__ __forceinline__ void advance(const DevModel& m, const HitRZ& h, double x0, double y0, double z0,
                                                 double u, double v, double w, Cell c,
                                                 double& x1, double& y1, double& z1, Cell& nxt) {
    const int ri0 = c.ri, tj0 = c.zj, pk0 = c.k;
    const int atj = tj0 < 0 ? -tj0 : tj0;
    exit_point(h, x0, y0, z0, u, v, w, x1, y1, z1);
    if (h.which == 0) {
      nxt.ri = ri0 + h.d_rad; nxt.zj = tj0; nxt.k = pk0;
      if (ri0 == 0) theta_phi_index(m, x1, y1, z1, nxt.zj, nxt.k);
      if (nxt.ri == 0) { nxt.zj = 1; nxt.k = 1; }
    } else if (h.which == 1) {
      nxt.ri = ri0; nxt.zj = atj + h.d_j;
      if (L3D) { if (z1 < 0) nxt.zj = -nxt.zj; }
      nxt.k = pk0;
    } else {
      nxt.ri = ri0; nxt.zj = tj0;
      int k1 = pk0 + h.d_phi;
      if (k1 == 0) k1 = m.n_az;
      if (k1 == m.n_az + 1) k1 = 1;
      nxt.k = k1;
    }
    if (z1 == 0.0) z1 = MCB_GRID_PREC;
  }

  // distance_to_closest_wall_sph (spherical_grid.f90:451-499).  The reference forms the theta-wall distances with
  // cos_phi_lim(thetaj0) (:471-472): the azimuthal table (size n_az, zero in 2D) indexed with the theta index, an
  // out-of-bounds read in dead code.  The distance from (rcyl, |z|) to the cone of theta_lim(j) is
  // |rcyl sin(theta_lim(j)) - |z| cos(theta_lim(j))|: cos(theta_lim) (tabulated on the host) replaces the mis-indexed table.
  static __device__ double closest_wall(const DevModel& m, Cell c, double x, double y, double z) {
    const int ri0 = c.ri, tj = c.zj < 0 ? -c.zj : c.zj, k0 = c.k;
    const double r2_cyl = x * x + y * y;
    const double rcyl = sqrt(r2_cyl);
    const double r = sqrt(r2_cyl + z * z);
    const double s1 = __ldg(m.r_lim + ri0) - r;
    const double s2 = r - __ldg(m.r_lim + ri0 - 1);
    const double z0 = fabs(z);
    const double s3 = fabs(rcyl * __ldg(m.w_lim + tj) - z0 * __ldg(m.cos_theta_lim + tj));
    const double s4 = fabs(rcyl * __ldg(m.w_lim + tj - 1) - z0 * __ldg(m.cos_theta_lim + tj - 1));
    double s = fmin(fmin(s1, s2), fmin(s3, s4));
    if (L3D) {
      const int km = (k0 - 1 >= 1) ? k0 - 1 : m.n_az;
      const double s5 = fabs(x * sin_phi_wall(m, k0) - y * __ldg(m.cos_phi_lim + k0 - 1));
      const double s6 = fabs(x * sin_phi_wall(m, km) - y * __ldg(m.cos_phi_lim + km - 1));
      s = fmin(s, fmin(s5, s6));
    }
    return s;
  }

  static __device__ __forceinline__ double cross(const DevModel& m, DirInv d, double x0, double y0, double z0,
                                                 double u, double v, double w, Cell c, Cell prev,
                                                 double& x1, double& y1, double& z1, Cell& nxt,
                                                 double& l_contrib, double& l_void) {
    HitRZ h = distance(m, d, x0, y0, z0, u, v, w, c, prev);
    advance(m, h, x0, y0, z0, u, v, w, c, x1, y1, z1, nxt);
    l_contrib = h.l; l_void = 0.0;
    return h.l;
  }

  static __device__ bool move_to_grid(const DevModel& m, double& x, double& y, double& z, double u, double v, double w, Cell& c) {
    const double correct_moins = 1.0 - 1.0e-10;
    double r0_2 = x * x + y * y + z * z;
    double b = (x * u + y * v + z * w);
    double cc = (r0_2 - r_lim_2<SM>(m, m.n_rad) * correct_moins);
    double delta = b * b - cc;
    if (delta < 0.0) return false;
    double s1 = -b - sqrt(delta);
    double x1 = x + s1 * u, y1 = y + s1 * v, z1 = z + s1 * w;
    c = index(m, x1, y1, z1);
    x = x1; y = y1; z = z1;
    return true;
  }

  static __device__ void pos_em_cell(const DevModel& m, Cell c, float rand1, float rand2, float rand3, double& x, double& y, double& z) {
    const int ri = c.ri, tj = c.zj, atj = tj < 0 ? -tj : tj;
    double r3a = __ldg(m.r_lim_3 + ri - 1), r3b = __ldg(m.r_lim_3 + ri);
    double r = pow(r3a + rand1 * (r3b - r3a), 1.0 / 3.0);
    double ta = __ldg(m.theta_lim + atj - 1), tb = __ldg(m.theta_lim + atj), theta;
    if (L3D) theta = ta + rand2 * (tb - ta);
    else theta = (rand2 > 0.5) ? ta + (2.0 * (rand2 - 0.5)) * (tb - ta) : -(ta + (2.0 * rand2) * (tb - ta));
    double phi = 2.0 * MCB_PI * ((double)(float)c.k - 1.0 + rand3) / (double)(float)m.n_az;
    double st, ct, sp, cp;
    sincos(theta, &st, &ct); sincos(phi, &sp, &cp);
    z = r * st;
    double rc = r * ct;
    x = rc * cp; y = rc * sp;
  }
};

}  // namespace mcb
