/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * A from-scratch CPU restatement (C++17 + OpenMP) of the reference algorithm of
 * MCFOST's Monte Carlo photon-packet loop (cpinte/mcfost 4.1.13).  It exists to
 * CHECK the CUDA path; it is never linked into, imported by, or called from the
 * product (mcfost_b200/).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it.
 *
 * PARITY STATUS: **parity unpinned by the reference.**  The reference cannot be
 * compiled or run in this environment (no Fortran compiler, link deps cfitsio /
 * SPRNG / voro++ absent, MCFOST_UTILS data absent) and its test-suite holds no
 * golden vectors for these routines.  The oracle is therefore
 * pinned only by (a) line-by-line review against the cited Fortran, (b) analytic
 * properties checked in tests/ (chord lengths, point-location round trips,
 * energy conservation, grey radiative-equilibrium temperature) and (c) its own
 * golden vectors committed under tests/golden/ (regression pins).
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/src).  Precision mirrors the Fortran: `real` = float,
 * `real(kind=dp)` = double; fp32 literals that the Fortran promotes to dp are
 * written (double)1.0e-10f etc.  Build the parity flavour with
 * -O2 -fno-fast-math -ffp-contract=off (see oracle/Makefile).
 *
 * Indices: all cell / wavelength / temperature indices in this file are the
 * reference's 1-based values; accessors subtract the offset.
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/mcfost_b200.h"
#include "philox.h"

namespace {

// ---- constants.f90:8-14, 151-159; cylindrical_grid.f90:16; spherical_grid.f90:18-19
const double pi = 3.141592653589793238462643383279502884197;
const double two_pi = 2.0 * pi;
const double inv_two_pi = 1.0 / two_pi;
const double half_pi = 0.5 * pi;
const double one_third = 1.0 / 3.0;
const float  tiny_real = std::numeric_limits<float>::min();
const float  huge_real = std::numeric_limits<float>::max();
const double tiny_dp = std::numeric_limits<double>::min();
const double huge_dp = std::numeric_limits<double>::max();
const float  tiny_real_x1e6 = tiny_real * 1.0e6f;
const float  max_int = (float)2147483647 * (1.0f - 1.0e-5f);
const double grid_prec = 1.0e-14;
const double prec_grille_sph = 1.0e-7;
const double hp = 6.626070040e-34, kb = 1.38064852e-23, c_light = 299792458.0;      // constants.f90:21-23
const double AU_to_cm = 149597870700.0 * 100.0;     // constants.f90:62-65
const double mum_to_cm = 1.0e-4;                    // constants.f90:73
const double AU_to_cm_mum2 = AU_to_cm * (mum_to_cm * mum_to_cm);   // AU_to_cm * mum_to_cm**2
const int nang_scatt = MCB_NANG_SCATT;
const int n_az_rt = MCB_N_AZ_RT;

// Fortran MODULO for reals (result has the sign of p)
inline double fmodulo(double a, double p) {
  double r = std::fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  return r;
}
// Fortran NINT (round half away from zero)
inline int f_nint(double x) { return (int)std::lround(x); }

struct Packet {   // the local variables of mc_photon_loop, dust_transfer.f90:456-467
  double x, y, z, u, v, w;
  double S[4];
  int icell, lambda;
  bool flag_star, flag_ISM, flag_scatt, alive, lintersect;
};

struct ThreadTallies {
  std::vector<double> xKJ_abs;      // (n_cells)
  std::vector<double> xJ_abs;       // (n_cells, n_lambda)
  std::vector<int>    xT_ech;       // (n_cells)
  std::vector<int>    xT_ech_1grain, xT_ech_1grain_nRE;   // (grains of the regime, n_cells) thermal_emission.f90:50
  double E_abs_nRE = 0.0;           // omp reduction variable, dust_transfer.f90:489
  std::vector<double> stokes_map, star_origin, disk_origin;   // output.f90:26-37
  std::vector<double> xN_abs;       // radiation_field.f90:23
  std::vector<double> n_phot_envoyes;
  std::vector<double> sed[9];       // sed, q, u, v, n_phot_sed, star, star_scat, disk, disk_scat
  std::vector<float>  xI_scatt;
  std::vector<float>  I_spec, I_spec_star;      // rt2, dust_ray_tracing.f90:44-45
  // per-thread scratch of angles_scatt_rt1 (dust_ray_tracing.f90: itheta_rt1 etc.)
  std::vector<int> itheta_rt1;
  std::vector<double> cos_omega_rt1, sin_omega_rt1;
  double stats[12];
};

struct Oracle {
  mcb_grid g{};
  mcb_opacity o{};
  mcb_emission e{};
  mcb_run_params r{};
  mcb_grains gr{};
  bool has_grid = false, has_op = false, has_em = false, has_gr = false;
  char err[256] = {0};

  // ---- build_cylindrical_cell_mapping, cylindrical_grid.f90:45-179
  int j_start = 1, jstart2 = 0, jend2 = 0, nj2 = 0, ntot2 = 0;
  std::vector<int> cmap, cmap_i, cmap_j, cmap_k, lexit;
  std::vector<int> dark;   // l_dark_zone(1:n_cells), index 0 unused
  int nb_proc = 1;
  std::vector<ThreadTallies> T;
  int N_type_flux = 1, n_Stokes = 1;
  double* ev_out = nullptr;      // optional per-packet event counts (instrumentation)
  uint8_t* ev_log = nullptr; int64_t ev_log_cap = 0, ev_log_n = 0;   // optional event trace (single thread): 1 step, 2 scatter, 3 absorb, 0 packet end
  inline void logev(uint8_t c) { if (ev_log && ev_log_n < ev_log_cap) ev_log[ev_log_n++] = c; }

  bool lvariable_dust() const { return o.p_n_cells != 1; }
  bool lVoronoi() const { return g.kind == MCB_GRID_VORONOI; }

  // ---------------- accessors (1-based Fortran indices) -----------------
  inline int& cell_map(int i, int j, int k) { return cmap[(size_t)i + (size_t)(g.n_rad + 2) * ((size_t)(j - jstart2) + (size_t)nj2 * (k - 1))]; }
  inline double r_lim(int i) const { return g.r_lim[i]; }
  inline double r_lim_2(int i) const { return g.r_lim_2[i]; }
  inline double r_lim_3(int i) const { return g.r_lim_3[i]; }
  inline double z_lim(int i, int j) const { return g.z_lim[(size_t)(i - 1) + (size_t)g.n_rad * (j - 1)]; }
  inline double zmax(int i) const { return g.zmax[i - 1]; }
  inline double tan_theta_lim(int j) const { return g.tan_theta_lim[j]; }
  inline double theta_lim(int j) const { return g.theta_lim[j]; }
  inline double tan_phi_lim(int k) const { return g.tan_phi_lim[k - 1]; }
  inline double volume(int ic) const { return g.volume[ic - 1]; }
  inline double kappa(int pc, int l) const { return o.kappa[(size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1)]; }
  inline double kappa_abs_LTE(int pc, int l) const { return o.kappa_abs_LTE[(size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1)]; }
  inline double kappa_factor(int ic) const { return o.kappa_factor[ic - 1]; }
  inline float tab_albedo_pos(int pc, int l) const { return o.tab_albedo_pos[(size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1)]; }
  inline float tab_g_pos(int pc, int l) const { return o.tab_g_pos[(size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1)]; }
  inline size_t pos_idx(int it, int pc, int pl) const { return (size_t)it + (size_t)(nang_scatt + 1) * ((size_t)(pc - 1) + (size_t)o.p_n_cells * (pl - 1)); }
  inline double log_Qcool(int t, int pc) const { return o.log_Qcool_minus_extra_heating[(size_t)(t - 1) + (size_t)o.n_T * (pc - 1)]; }
  inline double kdB_dT_CDF(int l, int t, int pc) const { return o.kdB_dT_CDF[(size_t)(l - 1) + (size_t)o.n_lambda * ((size_t)(t - 1) + (size_t)o.n_T * (pc - 1))]; }
  inline float tab_Temp(int t) const { return o.tab_Temp[t - 1]; }
  // per-grain tables (mcb_grains)
  inline size_t gl_idx(int k, int l) const { return (size_t)(k - 1) + (size_t)gr.n_grains_tot * (l - 1); }
  inline size_t cl_idx(int ic, int l) const { return (size_t)(ic - 1) + (size_t)g.n_cells * (l - 1); }
  inline double dust_density_o_n_grains(int pk, int ic) const { return gr.dust_density_o_n_grains[(size_t)(pk - 1) + (size_t)gr.n_dens * (ic - 1)]; }
  inline int nk_nLTE() const { return gr.grain_RE_nLTE_end - gr.grain_RE_nLTE_start + 1; }
  inline int nk_nRE() const { return gr.grain_nRE_end - gr.grain_nRE_start + 1; }
  inline double kabs_nLTE_CDF(int k, int ic, int l) const { return gr.kabs_nLTE_CDF[(size_t)(k - gr.grain_RE_nLTE_start + 1) + (size_t)(nk_nLTE() + 1) * cl_idx(ic, l)]; }
  inline double log_E_em_1grain(int k, int t) const { return gr.log_E_em_1grain[(size_t)(k - gr.grain_RE_nLTE_start) + (size_t)nk_nLTE() * (t - 1)]; }
  inline double kdB_dT_1grain_nLTE_CDF(int l, int k, int t) const { return gr.kdB_dT_1grain_nLTE_CDF[(size_t)(l - 1) + (size_t)o.n_lambda * ((size_t)(k - gr.grain_RE_nLTE_start) + (size_t)nk_nLTE() * (t - 1))]; }
  inline double log_E_em_1grain_nRE(int k, int t) const { return gr.log_E_em_1grain_nRE[(size_t)(k - gr.grain_nRE_start) + (size_t)nk_nRE() * (t - 1)]; }
  inline double kdB_dT_1grain_nRE_CDF(int l, int k, int t) const { return gr.kdB_dT_1grain_nRE_CDF[(size_t)(l - 1) + (size_t)o.n_lambda * ((size_t)(k - gr.grain_nRE_start) + (size_t)nk_nRE() * (t - 1))]; }
  inline bool l_RE(int k, int ic) const { return gr.l_RE[(size_t)(k - gr.grain_nRE_start) + (size_t)nk_nRE() * (ic - 1)] != 0; }
  inline double kappa_abs_nLTE(int pc, int l) const { return gr.kappa_abs_nLTE[(size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1)]; }
  inline float prob_s11(int l, int k, int a) const { return gr.prob_s11[(size_t)(l - 1) + (size_t)o.n_lambda * ((size_t)(k - 1) + (size_t)gr.n_grains_tot * a)]; }
  inline size_t s_idx(int a, int k, int l) const { return (size_t)a + (size_t)(nang_scatt + 1) * gl_idx(k, l); }
  inline double ksca_CDF(int k, int pc, int l) const { return gr.ksca_CDF[(size_t)k + (size_t)(gr.n_grains_tot + 1) * ((size_t)(pc - 1) + (size_t)o.p_n_cells * (l - 1))]; }
  inline double prob_E_cell(int k, int l) const { return e.prob_E_cell[(size_t)k + (size_t)(g.n_cells + 1) * (l - 1)]; }
  inline float CDF_E_star(int l, int k) const { return e.CDF_E_star[(size_t)(l - 1) + (size_t)o.n_lambda * k]; }
  inline double star_x(int i) const { return g.star_xyzr[4 * (i - 1) + 0]; }
  inline double star_y(int i) const { return g.star_xyzr[4 * (i - 1) + 1]; }
  inline double star_z(int i) const { return g.star_xyzr[4 * (i - 1) + 2]; }
  inline double star_r(int i) const { return g.star_xyzr[4 * (i - 1) + 3]; }

  // =====================================================================
  // cylindrical_grid.f90:45-179  build_cylindrical_cell_mapping (also sph)
  // =====================================================================
  int build_cell_mapping() {
    const int n_rad = g.n_rad, nz = g.nz, n_az = g.n_az;
    j_start = g.l3D ? -nz : 1;                       // grid.f90:279-283
    int istart = 1, iend = n_rad, jstart = j_start, jend = nz, kstart = 1, kend = n_az;
    int ntot;
    if (j_start < 0) ntot = (iend - istart + 1) * (jend - jstart) * (kend - kstart + 1);
    else ntot = (iend - istart + 1) * (jend - jstart + 1) * (kend - kstart + 1);
    if (ntot != g.n_cells) { snprintf(err, sizeof err, "cell mapping: ntot=%d should be %d", ntot, g.n_cells); return MCB_ERR_BAD_ARG; }
    int istart2 = 0, iend2 = n_rad + 1;
    jstart2 = std::min(1, j_start) - 1;
    jend2 = nz + 1;
    int kstart2 = 1, kend2 = n_az;
    if (jstart2 < 0) ntot2 = (iend2 - istart2 + 1) * (jend2 - jstart2) * (kend2 - kstart2 + 1);
    else ntot2 = (iend2 - istart2 + 1) * (jend2 - jstart2 + 1) * (kend2 - kstart2 + 1);
    nj2 = jend2 - jstart2 + 1;
    cmap.assign((size_t)(n_rad + 2) * nj2 * n_az, 0);
    cmap_i.assign(ntot2 + 1, 0); cmap_j.assign(ntot2 + 1, 0); cmap_k.assign(ntot2 + 1, 0);
    lexit.assign(ntot2 + 1, 0);
    int icell = 0;
    for (int k = kstart; k <= kend; ++k)
      for (int j = j_start; j <= jend; ++j) {
        if (j == 0) continue;
        for (int i = istart; i <= iend; ++i) {
          ++icell;
          cmap_i[icell] = i; cmap_j[icell] = j; cmap_k[icell] = k;
          cell_map(i, j, k) = icell;
        }
      }
    if (icell != ntot) { snprintf(err, sizeof err, "cell mapping: missing real cells"); return MCB_ERR_BAD_ARG; }
    // virtual cells, :115-167
    for (int k = kstart; k <= kend; ++k)
      for (int j = jstart2; j <= jend2; j += (jend2 - jstart2))
        for (int i = istart2; i <= iend2; ++i) {
          ++icell;
          if (icell > ntot2) { snprintf(err, sizeof err, "cell mapping overflow"); return MCB_ERR_BAD_ARG; }
          if (std::abs(j) == jend2) lexit[icell] = 2;
          if (i == iend2) lexit[icell] = 1;
          cmap_i[icell] = i; cmap_j[icell] = j; cmap_k[icell] = k;
          cell_map(i, j, k) = icell;
        }
    for (int k = kstart; k <= kend; ++k)
      for (int j = jstart; j <= jend; ++j) {
        if (j == 0) continue;
        for (int i = istart2; i <= iend2; i += (iend2 - istart2)) {
          ++icell;
          if (icell > ntot2) { snprintf(err, sizeof err, "cell mapping overflow (2)"); return MCB_ERR_BAD_ARG; }
          if (i == iend2) lexit[icell] = 1;
          cmap_i[icell] = i; cmap_j[icell] = j; cmap_k[icell] = k;
          cell_map(i, j, k) = icell;
        }
      }
    if (icell != ntot2) { snprintf(err, sizeof err, "cell mapping: missing virtual cells %d %d", icell, ntot2); return MCB_ERR_BAD_ARG; }
    return MCB_OK;
  }

  // =====================================================================
  // cylindrical_grid.f90:680-704  test_exit_grid_cyl
  // =====================================================================
  bool test_exit_grid_cyl(int icell, double, double, double z) const {
    if (icell <= g.n_cells) return false;
    if (lexit[icell] == 0) return false;
    else if (lexit[icell] == 1) return true;
    else return std::fabs(z) > g.zmaxmax;
  }

  // =====================================================================
  // cylindrical_grid.f90:833-890  index_cell_cyl
  // =====================================================================
  void index_cell_cyl(double xin, double yin, double zin, int& icell) {
    const int n_rad = g.n_rad, nz = g.nz, n_az = g.n_az;
    double r2 = xin * xin + yin * yin, phi;
    int ri, ri_min, ri_max, ri_out, zj_out, phik_out;
    if (r2 < r_lim_2(0)) { ri_out = 0; zj_out = 1; phik_out = 1; }
    else if (r2 > g.Rmax2) { ri_out = n_rad + 1; zj_out = 1; phik_out = 1; }
    else {
      ri_min = 0; ri_max = n_rad; ri = (ri_min + ri_max) / 2;
      while ((ri_max - ri_min) > 1) {
        if (r2 > r_lim_2(ri)) ri_min = ri; else ri_max = ri;
        ri = (ri_min + ri_max) / 2;
      }
      ri_out = ri + 1;
      // :868  fp32 cast then fp32 min
      zj_out = (int)std::floor(std::min((float)(std::fabs(zin) / zmax(ri_out) * nz), max_int)) + 1;
      if (g.l3D) {
        if (zj_out > nz) zj_out = nz + 1;
        if (zin < 0.0) zj_out = -zj_out;
        if (zin != 0.0) {
          phi = fmodulo(std::atan2(yin, xin), 2 * pi);
          phik_out = (int)std::floor(phi / (2 * pi) * (double)(float)n_az) + 1;
          if (phik_out == n_az + 1) phik_out = n_az;
        } else phik_out = 1;
      } else {
        if (zj_out > nz) zj_out = nz + 1;
        phik_out = 1;
      }
    }
    icell = cell_map(ri_out, zj_out, phik_out);
  }

  // =====================================================================
  // cylindrical_grid.f90:918-1175  cross_cylindrical_cell
  // =====================================================================
  void cross_cylindrical_cell(double x0, double y0, double z0, double u, double v, double w, int cell, int /*previous_cell*/,
                              double& x1, double& y1, double& z1, int& next_cell, double& l, double& l_contrib, double& l_void_before) {
    const int n_rad = g.n_rad, nz = g.nz, N_az = g.n_az;
    const bool l3D = g.l3D != 0;
    int ri0, zj0, k0, k0m1, delta_rad = 0, delta_zj = 0, delta_phi = 0, ri1, zj1, k1;
    double inv_a, a, b, c, s, rac, t, t_phi, delta, inv_w, r_2, den, tan_angle_lim, phi, delta_vol, zlim, dotprod;
    const double correct_moins = 1.0 - grid_prec, correct_plus = 1.0 + grid_prec;

    a = u * u + v * v;
    if (a > tiny_real) inv_a = 1.0 / a; else inv_a = huge_real;
    if (std::fabs(w) > tiny_real) inv_w = 1.0 / w; else inv_w = std::copysign(huge_dp, w);

    ri0 = cmap_i[cell]; zj0 = cmap_j[cell]; k0 = cmap_k[cell];   // cell2cylindrical :765

    r_2 = x0 * x0 + y0 * y0;
    b = (x0 * u + y0 * v) * inv_a;

    if (ri0 == 0) {
      c = (r_2 - r_lim_2(0)) * inv_a;
      delta = b * b - c;
      rac = std::sqrt(delta);
      s = (-b + rac) * correct_plus;
      t = huge_real;
      t_phi = huge_real;
      delta_rad = 1;
    } else {
      // 1) radial interface :973-1000
      dotprod = u * x0 + v * y0;
      if (dotprod < 0.0) {
        c = (r_2 - r_lim_2(ri0 - 1) * correct_moins) * inv_a;
        delta = b * b - c;
        if (delta < 0.0) {
          c = (r_2 - r_lim_2(ri0) * correct_plus) * inv_a;
          delta = std::max(b * b - c, 0.0);
          delta_rad = 1;
        } else delta_rad = -1;
      } else {
        c = (r_2 - r_lim_2(ri0) * correct_plus) * inv_a;
        delta = std::max(b * b - c, 0.0);
        delta_rad = 1;
      }
      rac = std::sqrt(delta);
      s = (-b - rac) * correct_plus;
      if (s < 0.0) s = (-b + rac) * correct_plus;
      else if (s == 0.0) s = grid_prec;

      // 2) vertical interface :1003-1055
      dotprod = w * z0;
      if (dotprod == 0.0) t = (double)1.0e10f;
      else {
        if (dotprod > 0.0) {
          if (std::abs(zj0) == nz + 1) { delta_zj = 0; zlim = std::copysign(1.0e10, z0); }
          else {
            zlim = std::copysign(z_lim(ri0, std::abs(zj0) + 1) * correct_plus, z0);
            delta_zj = 1;
            if (l3D && (z0 < 0.0)) delta_zj = -1;
          }
        } else {
          if (l3D) {
            if (z0 > 0.0) { zlim = z_lim(ri0, std::abs(zj0)) * correct_moins; delta_zj = -1; if (zj0 == 1) delta_zj = -2; }
            else { zlim = -z_lim(ri0, std::abs(zj0)) * correct_moins; delta_zj = 1; if (zj0 == -1) delta_zj = 2; }
          } else {
            if (zj0 == 1) {
              delta_zj = 1;
              if (z0 > 0.0) zlim = -z_lim(ri0, 2) * correct_moins; else zlim = z_lim(ri0, 2) * correct_moins;
            } else {
              if (z0 > 0.0) zlim = z_lim(ri0, zj0) * correct_moins; else zlim = -z_lim(ri0, zj0) * correct_moins;
              delta_zj = -1;
            }
          }
        }
        t = (zlim - z0) * inv_w;
        if (t < 0.0) t = grid_prec;
      }

      // 3) azimuthal interface :1058-1094
      if (l3D) {
        dotprod = x0 * v - y0 * u;
        if (std::fabs(dotprod) < (double)1.0e-10f) t_phi = (double)1.0e30f;
        else {
          if (dotprod > 0.0) { tan_angle_lim = tan_phi_lim(k0); delta_phi = 1; }
          else { k0m1 = k0 - 1; if (k0m1 == 0) k0m1 = N_az; tan_angle_lim = tan_phi_lim(k0m1); delta_phi = -1; }
          if (tan_angle_lim > 1.0e299) {
            if (std::fabs(u) > (double)1e-6f) t_phi = -x0 / u; else t_phi = (double)1.0e30f;
          } else {
            den = v - u * tan_angle_lim;
            if (std::fabs(den) > (double)1.0e-6f) t_phi = -(y0 - x0 * tan_angle_lim) / den; else t_phi = (double)1.0e30f;
          }
          if (t_phi < 0.0) t_phi = (double)1.0e30f;
        }
      } else t_phi = huge_real;
    }

    // 4) which interface :1098-1156
    if ((s < t) && (s < t_phi)) {
      l = s; delta_vol = s;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0 + delta_rad;
      if (ri1 == 0) { zj1 = 1; k1 = 1; }
      else {
        if (ri1 > n_rad) zj1 = zj0;
        else {
          zj1 = (int)std::floor(std::min((float)(std::fabs(z1) / zmax(ri1) * nz), max_int)) + 1;   // :1116 fp32
          if (zj1 > nz) zj1 = nz + 1;
          if (l3D && (z1 < 0.0)) zj1 = -zj1;
        }
        k1 = k0;
        if ((ri0 == 0) && l3D) {
          phi = fmodulo(std::atan2(y1, x1), 2 * pi);
          k1 = (int)std::floor(phi * inv_two_pi * (double)(float)N_az) + 1;
          if (k1 == N_az + 1) k1 = N_az;
        }
      }
    } else if (t < t_phi) {
      l = t; delta_vol = t;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0; zj1 = zj0 + delta_zj; k1 = k0;
    } else {
      l = t_phi; delta_vol = correct_plus * t_phi;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0;
      zj1 = (int)std::floor(std::fabs(z1) / zmax(ri1) * nz) + 1;        // :1150 fp64
      if (zj1 > nz) zj1 = nz + 1;
      if (z1 < 0.0) zj1 = -zj1;
      k1 = k0 + delta_phi;
      if (k1 == 0) k1 = N_az;
      if (k1 == N_az + 1) k1 = 1;
    }
    if (z1 == 0.0) { if (l3D) z1 = std::copysign(grid_prec, w); else z1 = grid_prec; }   // :1159-1165
    next_cell = cell_map(ri1, zj1, k1);
    l_contrib = l; l_void_before = 0.0;
  }

  // =====================================================================
  // cylindrical_grid.f90:1284-1411  move_to_grid_cyl
  // =====================================================================
  void move_to_grid_cyl(double& x, double& y, double& z, double u, double v, double w, int& icell, bool& lintersect) {
    double x0, y0, z0, z1, a, inv_a, r_2, b, c, delta, rac, s1, s2, dotprod, t1, t2, zlim, zlim2, delta_vol, inv_w;
    const double correct_moins = 1.0 - 1.0e-10;
    x0 = x; y0 = y; z0 = z;
    a = u * u + v * v;
    if (a > tiny_real) inv_a = 1.0 / a; else inv_a = huge_real;
    if (std::fabs(w) > tiny_real) inv_w = 1.0 / w; else inv_w = std::copysign(huge_dp, w);
    r_2 = x0 * x0 + y0 * y0;
    b = (x0 * u + y0 * v) * inv_a;
    c = (r_2 - r_lim_2(g.n_rad) * correct_moins) * inv_a;
    delta = b * b - c;
    if (delta < 0.0) { s1 = huge_real; s2 = huge_real; }
    else { rac = std::sqrt(delta); s1 = -b - rac; s2 = -b + rac; }
    dotprod = w * z0;
    if (std::fabs(dotprod) < tiny_real) { t1 = huge_real; t2 = huge_real; }
    else {
      if (z0 > 0.0) { zlim = g.zmaxmax * correct_moins; zlim2 = -g.zmaxmax * correct_moins; }
      else { zlim = -g.zmaxmax * correct_moins; zlim2 = g.zmaxmax * correct_moins; }
      t1 = (zlim - z0) * inv_w; t2 = (zlim2 - z0) * inv_w;
    }
    if (t1 > (double)1e20f) { if (s1 > (double)1e20f) { lintersect = false; return; } }
    if (t1 > s1) {
      if (t1 > s2) {
        delta_vol = s1;
        z1 = z0 + delta_vol * w;
        if (std::fabs(z1) > g.zmaxmax) { lintersect = false; return; }
        else lintersect = true;
      } else { lintersect = true; delta_vol = t1; }
    } else {
      if (t2 < s1) { lintersect = false; return; }
      else { lintersect = true; delta_vol = s1; }
    }
    x = x0 + delta_vol * u; y = y0 + delta_vol * v; z = z0 + delta_vol * w;
    index_cell_cyl(x, y, z, icell);
  }

  // =====================================================================
  // cylindrical_grid.f90:1415-1466  pos_em_cell_cyl
  // =====================================================================
  void pos_em_cell_cyl(int icell, float rand1, float rand2, float rand3, double& x, double& y, double& z) {
    int ri = cmap_i[icell], zj = cmap_j[icell], phik = cmap_k[icell];
    double r = std::sqrt(r_lim_2(ri - 1) + rand1 * (r_lim_2(ri) - r_lim_2(ri - 1)));
    if (g.l3D) {
      if (zj > 0) z = z_lim(ri, zj) + rand2 * (z_lim(ri, zj + 1) - z_lim(ri, zj));
      else z = -(z_lim(ri, -zj) + rand2 * (z_lim(ri, -zj + 1) - z_lim(ri, -zj)));
    } else {
      if (rand2 > 0.5) z = z_lim(ri, zj) + (2.0 * (rand2 - 0.5)) * (z_lim(ri, std::abs(zj) + 1) - z_lim(ri, zj));
      else z = -(z_lim(ri, zj) + (2.0 * rand2) * (z_lim(ri, zj + 1) - z_lim(ri, zj)));
    }
    double phi = 2.0 * pi * ((double)phik - 1.0 + rand3) / (double)g.n_az;
    x = r * std::cos(phi); y = r * std::sin(phi);
  }

  // =====================================================================
  // spherical_grid.f90:24-44  test_exit_grid_sph
  // =====================================================================
  bool test_exit_grid_sph(int icell) const {
    if (icell <= g.n_cells) return false;
    return lexit[icell] == 1;
  }

  // spherical_grid.f90:129-178  indice_cellule_sph_theta
  void indice_cellule_sph_theta(double xin, double yin, double zin, int& thetaj_out, int& phik_out) const {
    double r02 = xin * xin + yin * yin, tan_theta, phi;
    if (r02 > tiny_dp) tan_theta = std::fabs(zin) / std::sqrt(r02); else tan_theta = (double)1.0e30f;
    int thetaj_min = 0, thetaj_max = g.nz, thetaj = (thetaj_min + thetaj_max) / 2;
    while ((thetaj_max - thetaj_min) > 1) {
      if (tan_theta > tan_theta_lim(thetaj)) thetaj_min = thetaj; else thetaj_max = thetaj;
      thetaj = (thetaj_min + thetaj_max) / 2;
    }
    thetaj_out = thetaj + 1;
    if (g.l3D) {
      if (zin < 0) thetaj_out = -thetaj_out;
      if (zin != 0.0) {
        phi = fmodulo(std::atan2(yin, xin), 2 * pi);
        phik_out = (int)std::floor(phi / (2 * pi) * (double)(float)g.n_az) + 1;
        if (phik_out == g.n_az + 1) phik_out = g.n_az;
      } else phik_out = 1;
    } else phik_out = 1;
  }

  // =====================================================================
  // spherical_grid.f90:48-125  index_cell_sph
  // =====================================================================
  void index_cell_sph(double xin, double yin, double zin, int& icell) {
    double r02 = xin * xin + yin * yin, r2 = r02 + zin * zin;
    int ri, ri_min, ri_max, ri_out, thetaj_out, phik_out;
    if (r2 < r_lim_2(0)) { ri_out = 0; thetaj_out = 1; phik_out = 1; }
    else if (r2 > g.Rmax2) { ri_out = g.n_rad + 1; thetaj_out = 1; phik_out = 1; }
    else {
      ri_min = 0; ri_max = g.n_rad; ri = (ri_min + ri_max) / 2;
      while ((ri_max - ri_min) > 1) {
        if (r2 > r_lim_2(ri)) ri_min = ri; else ri_max = ri;
        ri = (ri_min + ri_max) / 2;
      }
      ri_out = ri + 1;
      indice_cellule_sph_theta(xin, yin, zin, thetaj_out, phik_out);   // identical inline code :85-118
    }
    icell = cell_map(ri_out, thetaj_out, phik_out);
  }

  // =====================================================================
  // spherical_grid.f90:182-446  cross_spherical_cell
  // =====================================================================
  void cross_spherical_cell(double x0, double y0, double z0, double u, double v, double w, int cell, int /*previous_cell*/,
                            double& x1, double& y1, double& z1, int& next_cell, double& l, double& l_contrib, double& l_void_before) {
    const int nz = g.nz, N_az = g.n_az;
    const bool l3D = g.l3D != 0;
    const double correct_moins = 1.0 - prec_grille_sph, correct_plus = 1.0 + prec_grille_sph, precision = 1.0e-15;
    double b, c, s, rac, t, delta, r0_2, r0_2_cyl, delta_vol, dotprod, t_phi, tan_angle_lim, den;
    int ri0, thetaj0, ri1, thetaj1, delta_rad, delta_theta = 0, phik0, phik1, delta_phi = 0, phik0m1;
    double a_theta, b_theta, c_theta, tan2, tan_angle_lim1, tan_angle_lim2, t1, t2, t1_1, t1_2, t2_1, t2_2, a_theta_m1;

    ri0 = cmap_i[cell]; thetaj0 = cmap_j[cell]; phik0 = cmap_k[cell];
    r0_2_cyl = x0 * x0 + y0 * y0;
    r0_2 = r0_2_cyl + z0 * z0;
    b = (x0 * u + y0 * v + z0 * w);

    if (ri0 == 0) {
      c = (r0_2 - r_lim_2(0) * correct_plus);
      delta = b * b - c;
      rac = std::sqrt(delta);
      s = (-b + rac) * correct_plus;
      t = huge_real; delta_rad = 1; t_phi = huge_real;
    } else {
      dotprod = b;
      if (dotprod < 0.0) {
        c = (r0_2 - r_lim_2(ri0 - 1) * correct_moins);
        delta = b * b - c;
        if (delta < 0.0) { c = (r0_2 - r_lim_2(ri0) * correct_plus); delta = std::max(b * b - c, 0.0); delta_rad = 1; }
        else delta_rad = -1;
      } else { c = (r0_2 - r_lim_2(ri0) * correct_plus); delta = std::max(b * b - c, 0.0); delta_rad = 1; }
      rac = std::sqrt(delta);
      s = -b - rac;
      if (s < 0.0) s = -b + rac; else if (s == 0.0) s = grid_prec;

      // 2) theta cones :263-341
      if (z0 >= 0.0) {
        tan_angle_lim1 = tan_theta_lim(std::abs(thetaj0)) * correct_plus;
        tan_angle_lim2 = tan_theta_lim(std::abs(thetaj0) - 1) * correct_moins;
      } else {
        tan_angle_lim1 = -tan_theta_lim(std::abs(thetaj0)) * correct_plus;
        tan_angle_lim2 = -tan_theta_lim(std::abs(thetaj0) - 1) * correct_moins;
      }
      tan2 = tan_angle_lim1 * tan_angle_lim1;
      a_theta = w * w - tan2 * (u * u + v * v);
      a_theta_m1 = 1.0 / a_theta;
      b_theta = w * z0 - tan2 * (x0 * u + y0 * v);
      c_theta = z0 * z0 - tan2 * (x0 * x0 + y0 * y0);
      delta = b_theta * b_theta - a_theta * c_theta;
      if (delta < 0.0) t1 = 1.0e30;
      else {
        rac = std::sqrt(delta);
        t1_1 = (-b_theta - rac) * a_theta_m1;
        t1_2 = (-b_theta + rac) * a_theta_m1;
        if (t1_1 <= precision) { if (t1_2 <= precision) t1 = 1.0e30; else t1 = t1_2; }
        else { if (t1_2 <= precision) t1 = t1_1; else t1 = std::min(t1_1, t1_2); }
      }
      tan2 = tan_angle_lim2 * tan_angle_lim2;
      a_theta = w * w - tan2 * (u * u + v * v);
      a_theta_m1 = 1.0 / a_theta;
      b_theta = w * z0 - tan2 * (x0 * u + y0 * v);
      c_theta = z0 * z0 - tan2 * (x0 * x0 + y0 * y0);
      delta = b_theta * b_theta - a_theta * c_theta;
      if (delta < 0.0) t2 = 1.0e30;
      else {
        rac = std::sqrt(delta);
        t2_1 = (-b_theta - rac) * a_theta_m1;
        t2_2 = (-b_theta + rac) * a_theta_m1;
        if (t2_1 <= precision) { if (t2_2 <= precision) t2 = 1.0e30; else t2 = t2_2; }
        else { if (t2_2 <= precision) t2 = t2_1; else t2 = std::min(t2_1, t2_2); }
      }
      if (t1 < t2) { t = t1; delta_theta = 1; if (std::abs(thetaj0) == nz) delta_theta = 0; }
      else { t = t2; delta_theta = -1; if (std::abs(thetaj0) == 1) delta_theta = 0; }

      // 3) azimuth :343-380
      if (l3D) {
        dotprod = x0 * v - y0 * u;
        if (std::fabs(dotprod) < (double)1.0e-10f) { t_phi = (double)1.0e30f; delta_phi = 0; }
        else {
          if (dotprod > 0.0) { tan_angle_lim = tan_phi_lim(phik0); delta_phi = 1; }
          else { phik0m1 = phik0 - 1; if (phik0m1 == 0) phik0m1 = N_az; tan_angle_lim = tan_phi_lim(phik0m1); delta_phi = -1; }
          if (tan_angle_lim > 1.0e299) t_phi = -x0 / u;
          else {
            den = v - u * tan_angle_lim;
            if (std::fabs(den) > (double)1.0e-6f) t_phi = -(y0 - x0 * tan_angle_lim) / den;
            else { t_phi = (double)1.0e30f; delta_phi = 0; }
          }
          if (t_phi < 0.0) { t_phi = (double)1.0e30f; delta_phi = 0; }
        }
      } else t_phi = huge_real;
    }

    if ((s < t) && (s < t_phi)) {
      l = s; delta_vol = s;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0 + delta_rad; thetaj1 = thetaj0; phik1 = phik0;
      if (ri0 == 0) indice_cellule_sph_theta(x1, y1, z1, thetaj1, phik1);
      if (ri1 == 0) { thetaj1 = 1; phik1 = 1; }
    } else if (t < t_phi) {
      l = t; delta_vol = t;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0; thetaj1 = std::abs(thetaj0) + delta_theta;
      if (l3D) { if (z1 < 0) thetaj1 = -thetaj1; }
      phik1 = phik0;
    } else {
      l = t_phi; delta_vol = correct_plus * t_phi;
      x1 = x0 + delta_vol * u; y1 = y0 + delta_vol * v; z1 = z0 + delta_vol * w;
      ri1 = ri0; thetaj1 = thetaj0; phik1 = phik0 + delta_phi;
      if (phik1 == 0) phik1 = N_az;
      if (phik1 == N_az + 1) phik1 = 1;
    }
    if (z1 == 0.0) z1 = grid_prec;
    next_cell = cell_map(ri1, thetaj1, phik1);
    l_contrib = l; l_void_before = 0.0;
  }

  // =====================================================================
  // spherical_grid.f90:562-615  move_to_grid_sph
  // =====================================================================
  void move_to_grid_sph(double& x, double& y, double& z, double u, double v, double w, int& icell, bool& lintersect) {
    const double correct_moins = 1.0 - 1.0e-10;
    double x0 = x, y0 = y, z0 = z;
    double r0_2 = x0 * x0 + y0 * y0 + z0 * z0;
    double b = (x0 * u + y0 * v + z0 * w);
    double c = (r0_2 - r_lim_2(g.n_rad) * correct_moins);
    double delta = b * b - c;
    if (delta < 0.0) { lintersect = false; icell = 0; return; }
    lintersect = true;
    double rac = std::sqrt(delta);
    double s1 = -b - rac;
    double delta_vol = s1;
    double x1 = x0 + delta_vol * u, y1 = y0 + delta_vol * v, z1 = z0 + delta_vol * w;
    index_cell_sph(x1, y1, z1, icell);
    x = x1; y = y1; z = z1;
  }

  // =====================================================================
  // spherical_grid.f90:619-699  pos_em_cell_sph
  // =====================================================================
  void pos_em_cell_sph(int icell, float rand1, float rand2, float rand3, double& x, double& y, double& z) {
    int ri = cmap_i[icell], thetaj = cmap_j[icell], phik = cmap_k[icell];
    double r = std::pow(r_lim_3(ri - 1) + rand1 * (r_lim_3(ri) - r_lim_3(ri - 1)), one_third);
    double theta;
    if (g.l3D) theta = theta_lim(std::abs(thetaj) - 1) + rand2 * (theta_lim(std::abs(thetaj)) - theta_lim(std::abs(thetaj) - 1));
    else {
      if (rand2 > 0.5) theta = theta_lim(thetaj - 1) + (2.0 * (rand2 - 0.5)) * (theta_lim(thetaj) - theta_lim(thetaj - 1));
      else theta = -(theta_lim(thetaj - 1) + (2.0 * rand2) * (theta_lim(thetaj) - theta_lim(thetaj - 1)));
    }
    // :656  real(phik) and real(n_az) are fp32, promoted
    double phi = 2.0 * pi * ((double)(float)phik - 1.0 + rand3) / (double)(float)g.n_az;
    z = r * std::sin(theta);
    double r_cos_theta = r * std::cos(theta);
    x = r_cos_theta * std::cos(phi); y = r_cos_theta * std::sin(phi);
  }

  // ---- grid.f90:16-22 procedure pointers ----------------------------------
  bool test_exit_grid(int icell, double x, double y, double z) const {
    if (g.kind == MCB_GRID_CYL) return test_exit_grid_cyl(icell, x, y, z);
    if (g.kind == MCB_GRID_SPH) return test_exit_grid_sph(icell);
    return icell < 0;   // Voronoi.f90:1446-1459
  }
  void index_cell(double x, double y, double z, int& icell) {
    if (g.kind == MCB_GRID_CYL) index_cell_cyl(x, y, z, icell);
    else if (g.kind == MCB_GRID_SPH) index_cell_sph(x, y, z, icell);
    else index_cell_voronoi(x, y, z, icell);
  }
  void cross_cell(double x0, double y0, double z0, double u, double v, double w, int cell, int previous_cell,
                  double& x1, double& y1, double& z1, int& next_cell, double& l, double& l_contrib, double& l_void_before) {
    if (g.kind == MCB_GRID_CYL) cross_cylindrical_cell(x0, y0, z0, u, v, w, cell, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
    else if (g.kind == MCB_GRID_SPH) cross_spherical_cell(x0, y0, z0, u, v, w, cell, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
    else cross_Voronoi_cell(x0, y0, z0, u, v, w, cell, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
  }
  void move_to_grid(double& x, double& y, double& z, double u, double v, double w, int& icell, bool& lintersect) {
    if (g.kind == MCB_GRID_CYL) move_to_grid_cyl(x, y, z, u, v, w, icell, lintersect);
    else if (g.kind == MCB_GRID_SPH) move_to_grid_sph(x, y, z, u, v, w, icell, lintersect);
    else move_to_grid_Voronoi(x, y, z, u, v, w, icell, lintersect);
  }
  void pos_em_cell(int icell, float r1, float r2, float r3, double& x, double& y, double& z) {
    if (g.kind == MCB_GRID_CYL) pos_em_cell_cyl(icell, r1, r2, r3, x, y, z);
    else if (g.kind == MCB_GRID_SPH) pos_em_cell_sph(icell, r1, r2, r3, x, y, z);
    else { // Voronoi.f90:1510-1543: emission from the cell centre (displacement line commented out, :1539)
      x = g.vor_xyz[3 * (size_t)(icell - 1) + 0]; y = g.vor_xyz[3 * (size_t)(icell - 1) + 1]; z = g.vor_xyz[3 * (size_t)(icell - 1) + 2];
    }
  }

  // =====================================================================
  // Voronoi.f90:1289-1317  distance_to_wall
  // =====================================================================
  double distance_to_wall(double x, double y, double z, double u, double v, double w, int iwall) const {
    double n[3], p[3], r[3] = {x, y, z}, k[3] = {u, v, w};
    for (int a = 0; a < 3; ++a) n[a] = g.wall_x[iwall - 1][a];
    for (int a = 0; a < 3; ++a) p[a] = (double)g.wall_x[iwall - 1][3] * std::fabs(n[a]);
    float den = (float)(n[0] * k[0] + n[1] * k[1] + n[2] * k[2]);       // `real :: den`
    if (std::fabs(den) > tiny_real)
      return (n[0] * (p[0] - r[0]) + n[1] * (p[1] - r[1]) + n[2] * (p[2] - r[2])) / (double)den;
    return (double)huge_real;
  }
  // Voronoi.f90:1463-1478  is_in_volume
  bool is_in_volume(double x, double y, double z) const {
    if ((x > g.wall_x[0][3]) && (x < g.wall_x[1][3]))
      if ((y > g.wall_x[2][3]) && (y < g.wall_x[3][3]))
        if ((z > g.wall_x[4][3]) && (z < g.wall_x[5][3])) return true;
    return false;
  }
  // Voronoi.f90:1548-1572  index_cell_voronoi (O(n) brute force, fp32 distances)
  void index_cell_voronoi(double xin, double yin, double zin, int& icell) const {
    float dist2_min = huge_real;
    for (int i = 1; i <= g.n_cells; ++i) {
      const double* c = g.vor_xyz + 3 * (size_t)(i - 1);
      float dist2 = (float)((c[0] - xin) * (c[0] - xin) + (c[1] - yin) * (c[1] - yin) + (c[2] - zin) * (c[2] - zin));
      if (dist2 < dist2_min) { icell = i; dist2_min = dist2; }
    }
  }
  // Voronoi.f90:1321-1375  distance_to_star
  double distance_to_star(double x, double y, double z, double u, double v, double w, int& i_star) const {
    double d = huge_dp;
    i_star = 0;
    for (int i = 1; i <= g.n_stars; ++i) {
      double dr[3] = {x - star_x(i), y - star_y(i), z - star_z(i)};
      double b = dr[0] * u + dr[1] * v + dr[2] * w;
      double c = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] - star_r(i) * star_r(i);
      double delta = b * b - c;
      if (delta >= 0.) {
        double rac = std::sqrt(delta), s1 = -b - rac;
        if (s1 < 0) { double s2 = -b + rac; if (s2 > 0) { d = 0.0; i_star = i; } }
        else if (s1 < d) { d = s1; i_star = i; }
      }
    }
    return d;
  }
  // =====================================================================
  // Voronoi.f90:839-992  cross_Voronoi_cell   (fp32 plane geometry!)
  // =====================================================================
  void cross_Voronoi_cell(double x, double y, double z, double u, double v, double w, int icell, int previous_cell,
                          double& x1, double& y1, double& z1, int& next_cell, double& s, double& s_contrib, double& s_void_before) {
    const double prec = (double)1e-5f;   // `real(kind=dp), parameter :: prec = 1e-5` : fp32 literal
    double s_tmp, den;
    float n[3], p[3], r[3], k[3], r_cell[3], r_neighbour[3];
    r[0] = (float)x; r[1] = (float)y; r[2] = (float)z;
    k[0] = (float)u; k[1] = (float)v; k[2] = (float)w;
    s = (double)1e30f;
    next_cell = 0;
    const double* cxyz = g.vor_xyz + 3 * (size_t)(icell - 1);
    for (int a = 0; a < 3; ++a) r_cell[a] = (float)cxyz[a];
    int ifirst = g.vor_first[icell - 1], ilast = g.vor_last[icell - 1];
    bool was_cut = g.vor_was_cut[icell - 1] != 0;
    double h = g.vor_h[icell - 1];
    bool is_a_star_neighbour = g.vor_is_star_neighbour[icell - 1] != 0;
    for (int i = ifirst; i <= ilast; ++i) {
      int id_n = g.neighbours_list[i - 1];
      if (id_n == previous_cell) continue;
      if (id_n > 0) {
        const double* nxyz = g.vor_xyz + 3 * (size_t)(id_n - 1);
        for (int a = 0; a < 3; ++a) r_neighbour[a] = (float)nxyz[a];     // Voronoi_xyz is a `real` copy (:61)
        for (int a = 0; a < 3; ++a) n[a] = r_neighbour[a] - r_cell[a];
        den = (double)(n[0] * k[0] + n[1] * k[1] + n[2] * k[2]);        // fp32 dot_product -> dp
        if (den <= 0.) continue;
        for (int a = 0; a < 3; ++a) p[a] = 0.5f * (r_neighbour[a] + r_cell[a]);
        float dot = n[0] * (p[0] - r[0]) + n[1] * (p[1] - r[1]) + n[2] * (p[2] - r[2]);
        s_tmp = (double)dot / den;
        if (s_tmp < 0.) s_tmp = (double)huge_real;
      } else {
        s_tmp = distance_to_wall(x, y, z, u, v, w, -id_n);
        if (s_tmp < 0.) s_tmp = (double)huge_real;
      }
      if (s_tmp < s) { s = s_tmp; next_cell = id_n; }
    }
    s = s * (1.0 + prec);
    x1 = x + u * s; y1 = y + v * s; z1 = z + w * s;
    if (next_cell == 0) {          // :925-936 rounding fallback
      x1 = x; y1 = y; z1 = z; s = 0.0;
      if (is_in_volume(x, y, z)) {
        index_cell_voronoi(x, y, z, next_cell);
        if (icell == next_cell) next_cell = -1;
      } else next_cell = -1;
    }
    if (was_cut) {                 // :938-976
      double dr[3] = {(double)r[0] - (double)r_cell[0], (double)r[1] - (double)r_cell[1], (double)r[2] - (double)r_cell[2]};
      // delta_r = r - r_cell is an fp32 subtraction assigned to dp
      for (int a = 0; a < 3; ++a) dr[a] = (double)(r[a] - r_cell[a]);
      double b = dr[0] * (double)k[0] + dr[1] * (double)k[1] + dr[2] * (double)k[2];
      double hc = h * g.cutting_distance_o_h;
      double c = dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] - hc * hc;
      double delta = b * b - c;
      if (delta < 0.) { s_void_before = s; s_contrib = 0.0; }
      else {
        double rac = std::sqrt(delta), s1 = -b - rac, s2 = -b + rac;
        if (s1 < 0) {
          if (s2 < 0) { s_void_before = s; s_contrib = 0.0; }
          else { s_void_before = 0.0; s_contrib = std::min(s2, s); }
        } else {
          if (s1 < s) { s_void_before = s1; s_contrib = std::min(s2, s) - s1; }
          else { s_void_before = s; s_contrib = 0.0; }
        }
      }
    } else { s_void_before = 0.0; s_contrib = s; }
    if (is_a_star_neighbour) {     // :978-988
      int i_star;
      double d_to_star = distance_to_star(x, y, z, u, v, w, i_star);
      if (i_star > 0) if (d_to_star < s) { s_contrib = d_to_star; next_cell = g.star_icell[i_star - 1]; }
    }
  }
  // =====================================================================
  // Voronoi.f90:1379-1442  move_to_grid_Voronoi.  The reference finds the entry
  // cell with a per-wall kd-tree (kdtree2, third party); the nearest seed is
  // what it returns, so the oracle uses the brute-force nearest seed.
  // =====================================================================
  void move_to_grid_Voronoi(double& x, double& y, double& z, double u, double v, double w, int& icell, bool& lintersect) {
    const double prec = 1.e-6;
    double s_walls[6]; int order[6];
    for (int iw = 1; iw <= 6; ++iw) {
      double l = distance_to_wall(x, y, z, u, v, w, iw);
      if (l >= 0) s_walls[iw - 1] = l * (1.0 + prec); else s_walls[iw - 1] = (double)huge_real;
      order[iw - 1] = iw;
    }
    std::stable_sort(order, order + 6, [&](int a, int b) { return s_walls[a - 1] < s_walls[b - 1]; });
    double xt = x, yt = y, zt = z; bool found = false;
    for (int i = 0; i < 6; ++i) {
      double l = s_walls[order[i] - 1];
      xt = x + l * u; yt = y + l * v; zt = z + l * w;
      if (is_in_volume(xt, yt, zt)) { found = true; break; }
    }
    if (!found) { icell = 0; lintersect = false; return; }
    lintersect = true;
    x = xt; y = yt; z = zt;
    index_cell_voronoi(x, y, z, icell);
  }

  // =====================================================================
  // stars.f90:812-884  intersect_stars
  // =====================================================================
  void intersect_stars(double x, double y, double z, double u, double v, double w, bool& lintersect_stars, int& i_star, int& icell_star) const {
    double d_to_star = huge_dp;
    i_star = 0;
    for (int i = 1; i <= g.n_stars; ++i) {
      double dr[3] = {x - star_x(i), y - star_y(i), z - star_z(i)};
      double b = dr[0] * u + dr[1] * v + dr[2] * w;
      double c = (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]) - star_r(i) * star_r(i);
      double delta = b * b - c;
      if (delta >= 0.) {
        double rac = std::sqrt(delta), s1 = -b - rac;
        if (s1 < 0) { double s2 = -b + rac; if (s2 > 0) { d_to_star = 0.0; i_star = i; } }
        else if (s1 < d_to_star) { d_to_star = s1; i_star = i; }
      }
    }
    lintersect_stars = (i_star > 0);
    icell_star = lintersect_stars ? g.star_icell[i_star - 1] : 0;
  }

  // =====================================================================
  // utils.f90:1636-1688  cdapres
  // =====================================================================
  static void cdapres(double cospsi, double phi, double u0, double v0, double w0, double& u1, double& v1, double& w1) {
    double cpsi = cospsi, spsi = std::sqrt(1.0 - cpsi * cpsi), sphi = std::sin(phi), cphi = std::cos(phi);
    double a = spsi * cphi, b = spsi * sphi;
    if (std::fabs(w0) <= (double)0.999999f) {
      double c = std::sqrt(1.0 - w0 * w0), cm1 = 1.0 / c, aw0 = a * w0;
      u1 = (aw0 * u0 - b * v0) * cm1 + cpsi * u0;
      v1 = (aw0 * v0 + b * u0) * cm1 + cpsi * v0;
      w1 = cpsi * w0 - a * c;
    } else { u1 = a; v1 = b; w1 = cpsi; }
  }
  // =====================================================================
  // utils.f90:553-599  rotation
  // =====================================================================
  static void rotation(double xinit, double yinit, double zinit, double u1, double v1, double w1, double& xfin, double& yfin, double& zfin) {
    double cost, sint, sing, prod, theta;
    if (w1 > 0.999999999) { cost = 1.0; sint = 0.0; sing = 0.0; }
    else {
      if (std::fabs(u1) < tiny_real) { cost = 0.0; sint = 1.0; sing = std::sqrt(1.0 - w1 * w1); }
      else { theta = std::atan2(v1, u1); cost = std::cos(theta); sint = std::sin(theta); sing = std::sqrt(1.0 - w1 * w1); }
    }
    prod = cost * xinit + sint * yinit;
    xfin = sing * prod + w1 * zinit;
    yfin = cost * yinit - sint * xinit;
    zfin = sing * zinit - w1 * prod;
  }
  // =====================================================================
  // random_numbers.f90:32-51  random_isotropic_direction
  // =====================================================================
  static void random_isotropic_direction(PacketRng& rng, double& u, double& v, double& w) {
    float rand = (float)rng_next(&rng);
    w = 2.0 * rand - 1.0;
    double uv = std::sqrt(1.0 - w * w);
    rand = (float)rng_next(&rng);
    double phi = pi * (2.0 * rand - 1.0);
    u = uv * std::cos(phi); v = uv * std::sin(phi);
  }

  // =====================================================================
  // thermal_emission.f90:364-400  select_wl_em
  // =====================================================================
  void select_wl_em(float rand, int& lambda) const {
    int kmin = 0, kmax = o.n_lambda, k = (kmin + kmax) / 2;
    while (e.spectre_emission_cumul[k] != (double)rand) {
      if (e.spectre_emission_cumul[k] < (double)rand) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
      if ((kmax - kmin) <= 1) break;
    }
    lambda = kmax;
  }
  // stars.f90:75-104  select_star   (note k=(kmax-kmin)/2 initial value, unused)
  void select_star(int lambda, float rand, int& n_star) const {
    int kmin = 0, kmax = g.n_stars, k = (kmax - kmin) / 2;
    while ((kmax - kmin) > 1) {
      if (CDF_E_star(lambda, k) < rand) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
    }
    n_star = kmax;
  }
  // thermal_emission.f90:2044-2074  select_cellule
  void select_cellule(int lambda, float rand, int& icell) const {
    int kmin = 0, kmax = g.n_cells, k = (kmin + kmax) / 2;
    while ((kmax - kmin) > 1) {
      if (prob_E_cell(k, lambda) < (double)rand) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
    }
    icell = kmax;
  }

  // =====================================================================
  // stars.f90:108-169  emit_packet_uniform_sphere
  // =====================================================================
  void emit_packet_uniform_sphere(int i_star, float rand1, float rand2, float rand3, float rand4,
                                  int& icell, double& x, double& y, double& z, double& u, double& v, double& w, bool& lintersect) {
    const double precision = (double)1e-6;
    z = 2.0 * rand1 - 1.0;
    double srw02 = std::sqrt(1.0 - z * z);
    double argmt = pi * (2.0 * rand2 - 1.0);
    x = srw02 * std::cos(argmt); y = srw02 * std::sin(argmt);
    double cospsi = (double)std::sqrt(rand3);     // sqrt(real) is an fp32 sqrt in Fortran, then promoted
    double phi = 2.0 * pi * rand4;
    cdapres(cospsi, phi, x, y, z, u, v, w);
    double r_star = star_r(i_star) * (1.0 + precision);
    x = x * r_star; y = y * r_star; z = z * r_star;
    x = x + star_x(i_star); y = y + star_y(i_star); z = z + star_z(i_star);
    if (lVoronoi()) icell = g.star_icell[i_star - 1];
    else index_cell(x, y, z, icell);
    if (g.star_out_model[i_star - 1]) move_to_grid(x, y, z, u, v, w, icell, lintersect);
    else lintersect = true;
  }
  // =====================================================================
  // stars.f90:728-787  emit_packet_ISM
  // =====================================================================
  void emit_packet_ISM(PacketRng& rng, int& icell, double& x, double& y, double& z, double& u, double& v, double& w, double* stokes, bool& lintersect) {
    stokes[0] = 1.; stokes[1] = stokes[2] = stokes[3] = 0.;
    float rand1 = (float)rng_next(&rng), rand2 = (float)rng_next(&rng);
    z = 2.0 * rand1 - 1.0;
    double srw02 = std::sqrt(1.0 - z * z);
    double argmt = pi * (2.0 * rand2 - 1.0);
    x = srw02 * std::cos(argmt); y = srw02 * std::sin(argmt);
    float rand3 = (float)rng_next(&rng), rand4 = (float)rng_next(&rng);
    double cospsi = (double)(-std::sqrt(rand3));
    double phi = 2.0 * pi * rand4;
    cdapres(cospsi, phi, x, y, z, u, v, w);
    double l = e.R_ISM;
    x = e.centre_ISM[0] + x * l; y = e.centre_ISM[1] + y * l; z = e.centre_ISM[2] + z * l;
    move_to_grid(x, y, z, u, v, w, icell, lintersect);
  }

  // =====================================================================
  // dust_transfer.f90:1047-1151  emit_packet
  // =====================================================================
  void emit_packet(PacketRng& rng, Packet& p) {
    p.lintersect = true;
    float rand = (float)rng_next(&rng);
    if ((double)rand <= e.frac_E_stars[p.lambda - 1]) {
      p.flag_star = true; p.flag_ISM = false;
      rand = (float)rng_next(&rng);
      int i_star; select_star(p.lambda, rand, i_star);
      rand = (float)rng_next(&rng);
      float rand2 = (float)rng_next(&rng), rand3 = (float)rng_next(&rng), rand4 = (float)rng_next(&rng);
      emit_packet_uniform_sphere(i_star, rand, rand2, rand3, rand4, p.icell, p.x, p.y, p.z, p.u, p.v, p.w, p.lintersect);
      p.S[0] = e.E_paquet; p.S[1] = p.S[2] = p.S[3] = 0.0;
      if (r.lspot) {                                                   // :1094-1119 (`real` locals)
        const float z_spot = (float)std::cos((double)(r.theta_spot / 180.0f) * pi);
        const float x_spot = (float)(std::sin((double)(r.theta_spot / 180.0f) * pi) * std::cos((double)(r.phi_spot / 180.0f) * pi));
        const float y_spot = (float)(std::sin((double)(r.theta_spot / 180.0f) * pi) * std::sin((double)(r.phi_spot / 180.0f) * pi));
        const float cos_thet_spot = std::sqrt(1.0f - r.surf_fraction_spot);
        if ((double)x_spot * p.x + (double)y_spot * p.y + (double)z_spot * p.z > (double)cos_thet_spot * star_r(1)) {
          const float hc_lk = (float)(hp * c_light / (r.tab_lambda[p.lambda - 1] * 1e-6 * kb));
          const float correct_spot = (float)((std::exp((double)hc_lk / r.star1_T) - 1) / (double)(std::exp(hc_lk / r.T_spot) - 1));
          for (int a = 0; a < 4; ++a) p.S[a] = p.S[a] * correct_spot;
        }
      }
    } else if ((double)rand <= e.frac_E_disk[p.lambda - 1]) {
      p.flag_star = false; p.flag_ISM = false;
      rand = (float)rng_next(&rng);
      select_cellule(p.lambda, rand, p.icell);
      rand = (float)rng_next(&rng);
      float rand2 = (float)rng_next(&rng), rand3 = (float)rng_next(&rng);
      pos_em_cell(p.icell, rand, rand2, rand3, p.x, p.y, p.z);
      random_isotropic_direction(rng, p.u, p.v, p.w);
      p.S[0] = e.E_paquet; p.S[1] = p.S[2] = p.S[3] = 0.0;
      if (r.lweight_emission) p.S[0] = p.S[0] * e.correct_E_emission[p.icell - 1];      // :1140-1142
    } else {
      p.flag_star = false; p.flag_ISM = true;
      emit_packet_ISM(rng, p.icell, p.x, p.y, p.z, p.u, p.v, p.w, p.S, p.lintersect);
    }
  }

  // =====================================================================
  // dust_ray_tracing.f90:409-476  angles_scatt_rt1
  // =====================================================================
  void angles_scatt_rt1(ThreadTallies& t, double u, double v, double w) const {
    for (int ibin = 1; ibin <= r.RT_n_incl; ++ibin)
      for (int iaz = 1; iaz <= r.RT_n_az; ++iaz) {
        size_t ix = (size_t)(ibin - 1) + (size_t)r.RT_n_incl * (iaz - 1);
        double ur = r.tab_u_rt[ix], vr = r.tab_v_rt[ix], wr = r.tab_w_rt[ibin - 1];
        float cos_scatt = (float)(ur * u + vr * v + wr * w);
        // :433  acos(real) fp32; real(nang_scatt)/pi -> dp
        int k = f_nint((double)(std::acos(cos_scatt) * (float)nang_scatt) / pi);
        if (k > nang_scatt) k = nang_scatt;
        if (k < 1) k = 1;
        t.itheta_rt1[ix] = k;
        if (r.lsepar_pola) {
          double v1pi, v1pj, v1pk, xnyp, costhet, theta, omega, cosw, sinw;
          rotation(u, v, w, -ur, -vr, -wr, v1pi, v1pj, v1pk);
          xnyp = std::sqrt(v1pk * v1pk + v1pj * v1pj);
          if (xnyp < (double)1e-10f) { xnyp = 0.0; costhet = 1.0; }
          else costhet = -1.0 * v1pj / xnyp;
          theta = std::acos(costhet);
          if (theta >= pi) theta = 0.0;
          theta = theta + half_pi;
          omega = 2.0 * theta;
          if (v1pk < 0.0) omega = -1.0 * omega;
          cosw = std::cos(omega); sinw = std::sin(omega);
          if (std::fabs(cosw) < (double)1e-06f) cosw = 0.0;
          if (std::fabs(sinw) < (double)1e-06f) sinw = 0.0;
          t.cos_omega_rt1[ix] = cosw; t.sin_omega_rt1[ix] = sinw;
        }
      }
  }
  inline size_t xI_idx(int phik, int psup, int itype, int iRT, int icell) const {
    return (size_t)(phik - 1) + (size_t)n_az_rt * ((size_t)(psup - 1) + 2 * ((size_t)(itype - 1) + (size_t)N_type_flux * ((size_t)(iRT - 1) + (size_t)(r.RT_n_incl * r.RT_n_az) * (size_t)(icell - 1))));
  }
  // dust_ray_tracing.f90:480-529  calc_xI_scatt
  void calc_xI_scatt(ThreadTallies& t, int p_lambda, int icell, int phik, int psup, double l, double stokes, bool flag_star) {
    int p_icell = lvariable_dust() ? icell : 1;
    for (int ibin = 1; ibin <= r.RT_n_incl; ++ibin)
      for (int iaz = 1; iaz <= r.RT_n_az; ++iaz) {
        size_t ix = (size_t)(ibin - 1) + (size_t)r.RT_n_incl * (iaz - 1);
        int it = t.itheta_rt1[ix];
        double flux = l * stokes * (double)o.tab_s11_pos[pos_idx(it, p_icell, p_lambda)];
        int iRT = ibin + r.RT_n_incl * (iaz - 1);
        float& a = t.xI_scatt[xI_idx(phik, psup, 1, iRT, icell)];
        a = (float)((double)a + flux);
        if (r.lsepar_contrib) {
          float& b = t.xI_scatt[xI_idx(phik, psup, n_Stokes + (flag_star ? 2 : 4), iRT, icell)];
          b = (float)((double)b + flux);
        }
      }
  }
  // dust_ray_tracing.f90:533-632  calc_xI_scatt_pola
  void calc_xI_scatt_pola(ThreadTallies& t, int p_lambda, int icell, int phik, int psup, double l, const double* stokes, bool flag_star) {
    int p_icell = lvariable_dust() ? icell : 1;
    for (int ibin = 1; ibin <= r.RT_n_incl; ++ibin)
      for (int iaz = 1; iaz <= r.RT_n_az; ++iaz) {
        size_t ix = (size_t)(ibin - 1) + (size_t)r.RT_n_incl * (iaz - 1);
        int it = t.itheta_rt1[ix];
        size_t q = pos_idx(it, p_icell, p_lambda);
        float s11 = o.tab_s11_pos[q];
        float s12 = -s11 * o.tab_s12_o_s11_pos[q];
        float s22 = s11 * o.tab_s22_o_s11_pos[q];
        float s33 = -s11 * o.tab_s33_o_s11_pos[q];
        float s34 = -s11 * o.tab_s34_o_s11_pos[q];
        float s44 = -s11 * o.tab_s44_o_s11_pos[q];
        double M11 = s11, M22 = s22, M12 = s12, M21 = s12, M33 = s33, M44 = s44, M34 = -s34, M43 = s34;
        double cosw = t.cos_omega_rt1[ix], sinw = t.sin_omega_rt1[ix];
        // RPO(2,2)=-cosw RPO(2,3)=-sinw RPO(3,2)=-sinw RPO(3,3)=cosw ; ROP(2,2)=cosw ROP(2,3)=-sinw ROP(3,2)=sinw ROP(3,3)=cosw
        double C[4], D[4], S[4];
        C[1] = cosw * stokes[1] + (-sinw) * stokes[2];
        C[2] = sinw * stokes[1] + cosw * stokes[2];
        C[0] = stokes[0]; C[3] = stokes[3];
        D[0] = M11 * C[0] + M12 * C[1];
        D[1] = M21 * C[0] + M22 * C[1];
        D[2] = M33 * C[2] + M34 * C[3];
        D[3] = M43 * C[2] + M44 * C[3];
        S[1] = (-cosw) * D[1] + (-sinw) * D[2];
        S[2] = (-sinw) * D[1] + cosw * D[2];
        S[0] = D[0]; S[3] = D[3];
        int iRT = ibin + r.RT_n_incl * (iaz - 1);
        for (int is = 1; is <= 4; ++is) {
          float& a = t.xI_scatt[xI_idx(phik, psup, is, iRT, icell)];
          a = (float)((double)a + l * S[is - 1]);
        }
        if (r.lsepar_contrib) {
          double flux = l * S[0];
          float& b = t.xI_scatt[xI_idx(phik, psup, flag_star ? 6 : 8, iRT, icell)];
          b = (float)((double)b + flux);
        }
      }
  }

  // =====================================================================
  // radiation_field.f90:31-135  save_radiation_field  (rt2 branch: unsupported)
  // =====================================================================
  void save_radiation_field(ThreadTallies& t, int lambda, int p_lambda, int icell, const double* Stokes, double l,
                            double x0, double y0, double z0, double x1, double y1, double z1, double u, double v, double w, bool flag_star, bool flag_direct_star) {
    int p_icell = lvariable_dust() ? icell : 1;
    if (r.letape_th) {
      t.xKJ_abs[icell - 1] += kappa_abs_LTE(p_icell, lambda) * l * Stokes[0];      // lRE_LTE
      if (r.lxJ_abs_step1) t.xJ_abs[(size_t)(icell - 1) + (size_t)g.n_cells * (lambda - 1)] += l * Stokes[0];
      if (r.lxN_abs) t.xN_abs[icell - 1] += 1.0;                                    // lmcfost_lib :53 (no wavelength dependence)
    } else {
      if (r.lxJ_abs) {
        t.xJ_abs[(size_t)(icell - 1) + (size_t)g.n_cells * (lambda - 1)] += l * Stokes[0];
        if (r.lxN_abs) t.xN_abs[(size_t)(icell - 1) + (size_t)g.n_cells * (lambda - 1)] += 1.0;      // lProDiMo :60
      }
      if (r.lscatt_ray_tracing1) {
        double xm = 0.5 * (x0 + x1), ym = 0.5 * (y0 + y1), zm = 0.5 * (z0 + z1);
        int phi_k, psup;
        if (g.l3D) { phi_k = 1; psup = 1; }
        else {
          double phi_pos = std::atan2(xm, ym);
          phi_k = (int)std::floor(fmodulo(phi_pos, two_pi) / two_pi * n_az_rt) + 1;
          if (phi_k > n_az_rt) phi_k = n_az_rt;
          psup = (zm > 0.0) ? 1 : 2;
        }
        if (r.lsepar_pola) calc_xI_scatt_pola(t, p_lambda, icell, phi_k, psup, l, Stokes, flag_star);
        else calc_xI_scatt(t, p_lambda, icell, phi_k, psup, l, Stokes[0], flag_star);
      } else if (r.lscatt_ray_tracing2) {       // :91-130, only 2D
        if (flag_direct_star) {
          float& a = t.I_spec_star[icell - 1];
          a = (float)((double)a + l * Stokes[0]);
        } else {
          const int n_phi_I = r.n_phi_I, n_theta_I = r.n_theta_I;
          double xm = 0.5 * (x0 + x1), ym = 0.5 * (y0 + y1), zm = 0.5 * (z0 + z1);
          double phi_pos = std::atan2(xm, ym);
          double phi_vol = std::atan2(-u, -v) + two_pi;
          int phi_I = (int)std::floor(fmodulo(phi_vol - phi_pos, two_pi) / two_pi * n_phi_I) + 1;
          if (phi_I > n_phi_I) phi_I = 1;
          int theta_I;
          if (zm > 0.0) theta_I = (int)std::floor(0.5 * (w + 1.0) * n_theta_I) + 1;
          else theta_I = (int)std::floor(0.5 * (-w + 1.0) * n_theta_I) + 1;
          if (theta_I > n_theta_I) theta_I = n_theta_I;
          auto at = [&](int itype) -> float& {
            return t.I_spec[(size_t)(itype - 1) + (size_t)N_type_flux * ((size_t)(theta_I - 1) + (size_t)n_theta_I * ((size_t)(phi_I - 1) + (size_t)n_phi_I * (size_t)(icell - 1)))];
          };
          for (int is = 1; is <= n_Stokes; ++is) { float& a = at(is); a = (float)((double)a + l * Stokes[is - 1]); }
          if (r.lsepar_contrib) { float& b = at(n_Stokes + (flag_star ? 2 : 4)); b = (float)((double)b + l * Stokes[0]); }
        }
      }
    }
  }

  // =====================================================================
  // optical_depth.f90:21-182  physical_length
  // =====================================================================
  void physical_length(ThreadTallies& t, int lambda, int p_lambda, const double* Stokes, int& icell,
                       double& xio, double& yio, double& zio, double& u, double& v, double& w,
                       bool flag_star, bool flag_direct_star, float extrin, float& ltot, bool& flag_sortie, bool& lpacket_alive,
                       bool tallies_on = true) {
    double x0, y0, z0, x1, y1, z1, x_old, y_old, z_old, extr;
    double l, tau, opacity, l_contrib, l_void_before;
    int icell_old, next_cell, previous_cell, icell_star, i_star, icell0;
    bool lcell_not_empty, lstop, lintersect_stars;
    lstop = false; flag_sortie = false;
    x0 = xio; y0 = yio; z0 = zio;
    x1 = xio; y1 = yio; z1 = zio;
    extr = extrin;
    next_cell = icell;
    icell0 = 0;
    ltot = 0.0f;
    if (tallies_on && (!r.letape_th) && r.lscatt_ray_tracing1) angles_scatt_rt1(t, u, v, w);
    intersect_stars(x0, y0, z0, u, v, w, lintersect_stars, i_star, icell_star);
    for (;;) {
      icell_old = icell0;
      x_old = x0; y_old = y0; z_old = z0;
      x0 = x1; y0 = y1; z0 = z1;
      previous_cell = icell0;
      icell0 = next_cell;
      if (test_exit_grid(icell0, x0, y0, z0)) { flag_sortie = true; return; }
      if (lintersect_stars) if (icell0 == icell_star) { lpacket_alive = false; flag_sortie = true; return; }
      if (icell0 <= g.n_cells) {
        lcell_not_empty = true;
        int p_icell = lvariable_dust() ? icell0 : 1;
        opacity = kappa(p_icell, lambda) * kappa_factor(icell0);
        if (dark[icell0]) {
          u = -u; v = -v; w = -w;
          icell = icell_old;
          xio = x_old; yio = y_old; zio = z_old;
          flag_sortie = false;
          t.stats[7] += 1;
          return;
        }
      } else { lcell_not_empty = false; opacity = 0.0; }
      cross_cell(x0, y0, z0, u, v, w, icell0, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
      t.stats[1] += 1; logev(1);
      tau = l_contrib * opacity;
      if (tau > extr) {
        lstop = true;
        l_contrib = l_contrib * (extr / tau);
        l = l_void_before + l_contrib;
        ltot = (float)((double)ltot + l);
      } else {
        extr = extr - tau;
        ltot = (float)((double)ltot + l);
      }
      if (lcell_not_empty && tallies_on)
        save_radiation_field(t, lambda, p_lambda, icell0, Stokes, l_contrib, x0, y0, z0, x1, y1, z1, u, v, w, flag_star, flag_direct_star);
      if (lstop) {
        flag_sortie = false;
        xio = x0 + l * u; yio = y0 + l * v; zio = z0 + l * w;
        icell = icell0;
        if (!lVoronoi()) if (g.l3D) if (g.kind == MCB_GRID_CYL) index_cell(xio, yio, zio, icell);
        return;
      }
    }
  }

  // =====================================================================
  // optical_depth.f90:248-324  optical_length_tot
  // =====================================================================
  void optical_length_tot(int lambda, int icell, double xi, double yi, double zi, double u, double v, double w,
                          float& tau_tot_out, double& lmin, double& lmax, int& n_steps) {
    double x0, y0, z0, x1, y1, z1, l, ltot, tau, opacity, tau_tot, l_contrib, l_void_before;
    int previous_cell, next_cell, icell0;
    x1 = xi; y1 = yi; z1 = zi;
    tau_tot = 0.0; lmin = 0.0; ltot = 0.0;
    next_cell = icell; icell0 = 0; n_steps = 0;
    for (;;) {
      previous_cell = icell0; icell0 = next_cell;
      x0 = x1; y0 = y1; z0 = z1;
      if (test_exit_grid(icell0, x0, y0, z0)) { tau_tot_out = (float)tau_tot; lmax = ltot; return; }
      if (icell0 <= g.n_cells) { int p_icell = lvariable_dust() ? icell0 : 1; opacity = kappa(p_icell, lambda) * kappa_factor(icell0); }
      else opacity = 0.0;
      cross_cell(x0, y0, z0, u, v, w, icell0, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
      ++n_steps;
      tau = l_contrib * opacity;
      tau_tot = tau_tot + tau;
      ltot = ltot + l;
      if (tau_tot < tiny_real) lmin = ltot;
    }
  }

  // =====================================================================
  // optical_depth.f90:328-415  compute_column: from the centre of every cell along 4 directions (towards the star at the
  // origin, +z, -z, radially outwards), the sum of l_contrib * factor, factor = kappa(lambda) * kappa_factor (type 2) or
  // the caller's per-cell array (types 1 and 3: CD_units * gas_density [* tab_abundance]).  The centres are the caller's
  // r_grid cos(phi_grid), r_grid sin(phi_grid), z_grid (or the Voronoi seeds), :362-370.
  // Deviation, on purpose: with lvariable_dust the reference takes p_icell from the PREVIOUS cell (:391 runs before :393;
  // index 0 on the first step, out of bounds); the cell itself is used here, as optical_length_tot does (:303-304).
  // =====================================================================
  void compute_column(int lambda, const double* factor, const double* cx, const double* cy, const double* cz, float* column) {
    const int n = g.n_cells;
    for (int direction = 1; direction <= 4; ++direction) {
      for (int icell = 1; icell <= n; ++icell) {
        double x1 = cx[icell - 1], y1 = cy[icell - 1], z1 = cz[icell - 1], u, v, w, norm;
        if (direction == 1) { norm = 1.0 / std::sqrt(x1 * x1 + y1 * y1 + z1 * z1); u = -x1 * norm; v = -y1 * norm; w = -z1 * norm; }
        else if (direction == 2) { u = 0.0; v = 0.0; w = 1.0; }
        else if (direction == 3) { u = 0.0; v = 0.0; w = -1.0; }
        else { u = x1; v = y1; w = 0.0; norm = 1.0 / std::sqrt(u * u + v * v); u = u * norm; v = v * norm; }
        int next_cell = icell, icell0 = 0, previous_cell;
        double sum = 0.0, x0, y0, z0, l, l_contrib, l_void_before;
        for (;;) {
          previous_cell = icell0; icell0 = next_cell;
          x0 = x1; y0 = y1; z0 = z1;
          if (test_exit_grid(icell0, x0, y0, z0)) break;
          cross_cell(x0, y0, z0, u, v, w, icell0, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
          if (icell0 <= n) {
            const int p_icell = lvariable_dust() ? icell0 : 1;
            const double f = factor ? factor[icell0 - 1] : kappa(p_icell, lambda) * kappa_factor(icell0);
            sum = sum + l_contrib * f;
          }
        }
        column[(size_t)(icell - 1) + (size_t)n * (direction - 1)] = (float)sum;
      }
    }
  }

  // =====================================================================
  // optical_depth.f90:1425-1651  define_dark_zone.  `real` sums and angles as in the Fortran; r_grid / z_grid are the
  // caller's cell centres; zj_sup / zj_inf are in-out (module arrays that keep their values between calls, mem.f90:162-170);
  // dust_sum = sum(dust_density_o_n_grains(:,icell)) when n_zones > 1, else null; regions as (iRmin, iRmax) pairs.
  // Kept as they are: the lower-half loop of the 3D branch marks nothing (`do jj=1,-1` runs zero times, :1607-1610) and
  // starts its rays at z0 = -z_grid of a cell below the midplane (:1597); a stale zj_inf = 0 would index row 0 (:1591),
  // those rows are skipped here.
  // =====================================================================
  void define_dark_zone(int lambda, float tau_max, const double* r_grid, const double* z_grid, int n_regions, const int* iRmin,
                        const int* iRmax, const double* dust_sum, int* dark_out, int* ri_in, int* ri_out, int* zj_sup, int* zj_inf,
                        int* l_is_dark) {
    const int nbre_angle = 11;
    const int n_rad = g.n_rad, nz = g.nz, n_az = g.n_az;
    const bool l3D = g.l3D != 0, lcyl = g.kind == MCB_GRID_CYL;
    auto ZS = [&](int i, int pk) -> int& { return zj_sup[(size_t)(i - 1) + (size_t)n_rad * (pk - 1)]; };
    auto ZI = [&](int i, int pk) -> int& { return zj_inf[(size_t)(i - 1) + (size_t)n_rad * (pk - 1)]; };
    auto kap = [&](int icell) { const int p_icell = lvariable_dust() ? icell : 1; return kappa(p_icell, lambda) * kappa_factor(icell); };
    for (int pk = 1; pk <= n_az; ++pk) {
      ri_in[pk - 1] = n_rad; ri_out[pk - 1] = 1;
      float total_sum = 0.0f;
      for (int i = 1; i <= n_rad; ++i) {
        const int icell = cell_map(i, 1, pk);
        total_sum = (float)((double)total_sum + kap(icell) * (r_lim(i) - r_lim(i - 1)));
        if (total_sum > tau_max) { ri_in[pk - 1] = i; break; }
      }
      total_sum = 0.0f;
      for (int i = n_rad; i >= 1; --i) {
        const int icell = cell_map(i, 1, pk);
        total_sum = (float)((double)total_sum + kap(icell) * (r_lim(i) - r_lim(i - 1)));
        if (total_sum > tau_max) { ri_out[pk - 1] = i; break; }
      }
      if (ri_out[pk - 1] == n_rad) ri_out[pk - 1] = n_rad - 1;
      if (lcyl) {
        for (int i = ri_in[pk - 1]; i <= ri_out[pk - 1]; ++i) {
          total_sum = 0.0f;
          for (int j = nz; j >= 1; --j) {
            const int icell = cell_map(i, j, pk);
            total_sum = (float)((double)total_sum + kap(icell) * (z_lim(i, j + 1) - z_lim(i, j)));
            if (total_sum > tau_max) { ZS(i, pk) = j; break; }
          }
        }
        if (l3D) {
          for (int i = ri_in[pk - 1]; i <= ri_out[pk - 1]; ++i) {
            total_sum = 0.0f;
            for (int j = -nz; j <= -1; ++j) {
              const int icell = cell_map(i, j, pk);
              total_sum = (float)((double)total_sum + kap(icell) * (z_lim(i, -j + 1) - z_lim(i, -j)));
              if (total_sum > tau_max) { ZI(i, pk) = j; break; }
            }
          }
        }
      } else {
        for (int i = 1; i <= n_rad; ++i) ZS(i, pk) = nz;
      }
    }
    *l_is_dark = 0;
    std::fill(dark.begin(), dark.end(), 0);
    ThreadTallies t; for (double& q : t.stats) q = 0;
    const double Stokes[4] = {0, 0, 0, 0};
    // one ray (optical_depth.f90:1531-1541); returns flag_sortie
    auto ray = [&](int icell, double x0, double y0, double z0, int n) {
      const float angle = (float)(pi * (double)(float)n / (double)(float)(nbre_angle + 1));      // pi * real(n) / real(nbre_angle+1) -> real
      double u0 = (double)std::cos(angle), v0 = 0.0, w0 = (double)std::sin(angle);
      int ic = icell; bool fs = false, alive = true; float lt = 0;
      physical_length(t, lambda, 1, Stokes, ic, x0, y0, z0, u0, v0, w0, false, false, tau_max, lt, fs, alive, false);
      return fs;
    };
    if (!l3D) {
      for (int i = std::max(ri_in[0], 2); i <= ri_out[0]; ++i) {
        bool done = false;
        for (int j = ZS(i, 1); j >= 1 && !done; --j) {
          const int icell = cell_map(i, j, 1);
          for (int n = 1; n <= nbre_angle; ++n) {
            if (!ray(icell, r_grid[icell - 1], 0.0, z_grid[icell - 1], n)) {
              for (int jj = 1; jj <= j; ++jj) dark[cell_map(i, jj, 1)] = 1;
              *l_is_dark = 1;
              done = true; break;
            }
          }
        }
      }
    } else {
      for (int pk = 1; pk <= n_az; ++pk) {
        const float phi = (float)(2.0 * pi * (double)((float)pk - 0.5f) / (double)(float)n_az);
        for (int i = std::max(ri_in[pk - 1], 2); i <= ri_out[pk - 1]; ++i) {
          bool done = false;
          for (int j = ZS(i, pk); j >= 1 && !done; --j) {
            const int icell = cell_map(i, j, pk);
            for (int n = 1; n <= nbre_angle; ++n) {
              const float r0 = (float)r_grid[icell - 1];
              const double x0 = (double)(r0 * std::cos(phi)), y0 = (double)(r0 * std::sin(phi)), z0 = z_grid[icell - 1];
              if (!ray(icell, x0, y0, z0, n)) {
                for (int jj = 1; jj <= j; ++jj) dark[cell_map(i, jj, pk)] = 1;
                done = true; break;
              }
            }
          }
        }
        for (int i = std::max(ri_in[pk - 1], 2); i <= ri_out[pk - 1]; ++i) {
          bool done = false;
          for (int j = ZI(i, pk); j <= -1 && !done; ++j) {
            if (j < -nz) continue;
            const int icell = cell_map(i, j, pk);
            for (int n = 1; n <= nbre_angle; ++n) {
              const float r0 = (float)r_grid[icell - 1];
              const double x0 = (double)(r0 * std::cos(phi)), y0 = (double)(r0 * std::sin(phi)), z0 = -z_grid[icell - 1];
              if (!ray(icell, x0, y0, z0, n)) { *l_is_dark = 1; done = true; break; }      // (`do jj=1,-1`: nothing is marked)
            }
          }
        }
      }
    }
    for (int pk = 1; pk <= n_az; ++pk) {
      for (int i = 1; i <= ri_in[pk - 1] - 1; ++i) ZS(i, pk) = ZS(ri_in[pk - 1], pk);
      for (int i = ri_out[pk - 1] + 1; i <= n_rad; ++i) ZS(i, pk) = ZS(ri_out[pk - 1], pk);
      if (l3D) {
        for (int i = 1; i <= ri_in[pk - 1] - 1; ++i) ZI(i, pk) = ZI(ri_in[pk - 1], pk);
        for (int i = ri_out[pk - 1] + 1; i <= n_rad; ++i) ZI(i, pk) = ZI(ri_out[pk - 1], pk);
      }
    }
    if (dust_sum) for (int icell = 1; icell <= g.n_cells; ++icell) if (dust_sum[icell - 1] < tiny_real) dark[icell] = 0;
    for (int q = 0; q < n_regions; ++q)
      for (int j = 1; j <= nz; ++j) { dark[cell_map(iRmin[q], j, 1)] = 0; dark[cell_map(iRmax[q], j, 1)] = 0; }
    for (int icell = 1; icell <= g.n_cells; ++icell) dark_out[icell - 1] = dark[icell];
  }

  // =====================================================================
  // thermal_emission.f90:404-644  init_reemission: Planck function and its temperature derivative per (lambda, T), the
  // cooling table log(Qcool - Qcool(T_min)) and the emission CDF of the LTE cells (high-memory branch, no extra heating),
  // and the per-grain tables of the nLTE / nRE grains.  `thermal_const` is a `real` parameter (constants.f90:24), 1.e-6
  // and 500.0 are `real` literals.
  // =====================================================================
  void planck_tables(const double* tab_lambda, const double* tab_delta_lambda, std::vector<double>& B, std::vector<double>& dB) const {
    const int n_lambda = o.n_lambda, n_T = o.n_T;
    const float thermal_const = (float)(299792458.0 * 6.626070040e-34 / 1.38064852e-23);
    B.assign((size_t)n_lambda * n_T, 0.0); dB.assign((size_t)n_lambda * n_T, 0.0);
    for (int t = 1; t <= n_T; ++t) {
      const double Temp = tab_Temp(t);
      const double cst = (double)thermal_const / Temp;
      for (int lambda = 1; lambda <= n_lambda; ++lambda) {
        const double wl = tab_lambda[lambda - 1] * (double)1.e-6f;
        const double delta_wl = tab_delta_lambda[lambda - 1] * (double)1.e-6f;
        const double cst_wl = cst / wl;
        if (cst_wl < 500.0) {
          const double coeff_exp = std::exp(cst_wl);
          const double wl2 = wl * wl, wl5 = (wl2 * wl2) * wl;      // wl**5 as the compiler expands a small integer power (square, square, multiply)
          const double b = 1.0 / (wl5 * (coeff_exp - 1.0)) * delta_wl;
          B[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)] = b;
          dB[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)] = b * cst_wl * coeff_exp / (coeff_exp - 1.0);
        }
      }
    }
  }
  void init_reemission(const double* tab_lambda, const double* tab_delta_lambda, double* logQ, double* cdf) const {
    const int n_lambda = o.n_lambda, n_T = o.n_T, pnc = o.p_n_cells;
    const double cst_E = 2.0 * 6.626070040e-34 * (299792458.0 * 299792458.0) * (4.0 * pi);
    std::vector<double> B, dB; planck_tables(tab_lambda, tab_delta_lambda, B, dB);
    std::vector<double> integ3(n_lambda + 1);
    for (int icell = 1; icell <= pnc; ++icell) {
      double Qcool0 = 0.0;
      for (int t = 1; t <= n_T; ++t) {
        double integ = 0.0;
        for (int lambda = 1; lambda <= n_lambda; ++lambda) integ = integ + kappa_abs_LTE(icell, lambda) * B[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)];
        const double Qcool = integ * cst_E;
        if (t == 1) Qcool0 = Qcool;
        const double q = Qcool - Qcool0;
        logQ[(size_t)(t - 1) + (size_t)n_T * (icell - 1)] = (q > tiny_dp) ? std::log(q) : -1000.0;
        integ3[0] = 0.0;
        for (int lambda = 1; lambda <= n_lambda; ++lambda) integ3[lambda] = integ3[lambda - 1] + kappa_abs_LTE(icell, lambda) * dB[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)];
        double* c = cdf + (size_t)n_lambda * ((size_t)(t - 1) + (size_t)n_T * (icell - 1));
        if (integ3[n_lambda] > tiny_dp) for (int lambda = 1; lambda <= n_lambda; ++lambda) c[lambda - 1] = integ3[lambda] / integ3[n_lambda];
        else for (int lambda = 1; lambda <= n_lambda; ++lambda) c[lambda - 1] = 0.0;      // (left as allocated by the reference: zero)
      }
    }
  }
  // grains k_start..k_end (1-based): log_E_em_1grain(k, T) (:551-567 nLTE, :585-603 nRE) and kdB_dT_1grain_*_CDF(lambda, k, T)
  // (:569-581, :605-618; the CDF starts at 0 for lambda = 1)
  void init_reemission_grains(const double* tab_lambda, const double* tab_delta_lambda, const float* C_abs_norm, int n_grains_tot, int k_start,
                              int k_end, double* logE, double* E_em, double* cdf) const {
    const int n_lambda = o.n_lambda, n_T = o.n_T, nk = k_end - k_start + 1;
    const double cst_E = 2.0 * 6.626070040e-34 * (299792458.0 * 299792458.0) * (4.0 * pi);
    std::vector<double> B, dB; planck_tables(tab_lambda, tab_delta_lambda, B, dB);
    std::vector<double> integ3(n_lambda + 1);
    for (int t = 1; t <= n_T; ++t)
      for (int k = k_start; k <= k_end; ++k) {
        auto ca = [&](int lambda) { return (double)C_abs_norm[(size_t)(k - 1) + (size_t)n_grains_tot * (lambda - 1)]; };
        double integ = 0.0;
        for (int lambda = 1; lambda <= n_lambda; ++lambda) integ = integ + ca(lambda) * B[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)];
        const size_t kt = (size_t)(k - k_start) + (size_t)nk * (t - 1);
        logE[kt] = (integ > tiny_dp) ? std::log(integ * cst_E) : -1000.0;
        if (E_em) E_em[kt] = integ * cst_E;
        integ3[1] = 0.0;
        for (int lambda = 2; lambda <= n_lambda; ++lambda) integ3[lambda] = integ3[lambda - 1] + ca(lambda) * dB[(size_t)(lambda - 1) + (size_t)n_lambda * (t - 1)];
        double* c = cdf + (size_t)n_lambda * kt;
        if (integ3[n_lambda] > tiny_dp) for (int lambda = 1; lambda <= n_lambda; ++lambda) c[lambda - 1] = integ3[lambda] / integ3[n_lambda];
        else for (int lambda = 1; lambda <= n_lambda; ++lambda) c[lambda - 1] = 0.0;
      }
  }

  // =====================================================================
  // dust_ray_tracing.f90:636-708  init_dust_source_fct1: the source function of ray-tracing method 1 from the scattered
  // specific intensity.  xI is xI_scatt(n_az_rt, n_theta_rt, N_type_flux, n_RT, n_cells) already summed over the threads
  // (`real`); storage extents are az_dim x th_dim (45 x 2), of which n_az_rt x n_theta_rt are used (1 x 1 on a 3D grid).
  // =====================================================================
  void init_dust_source_fct1(int lambda, int iRT, int n_RT, double photon_energy, const double* J_th, const float* xI, int az_dim, int th_dim,
                             int n_az_rt, int n_theta_rt, int N_type_flux, int n_Stokes, bool lsepar_pola, bool lsepar_contrib, double* eps) const {
    const size_t per_cell = (size_t)az_dim * th_dim * N_type_flux;
    std::fill(eps, eps + per_cell * (size_t)g.n_cells, 0.0);
    for (int icell = 1; icell <= g.n_cells; ++icell) {
      const int p_icell = lvariable_dust() ? icell : 1;
      const double factor = photon_energy / volume(icell) * n_az_rt * n_theta_rt;
      const double kappa_ext = kappa(p_icell, lambda) * kappa_factor(icell);
      const double kappa_sca = kappa_ext * tab_albedo_pos(p_icell, lambda);
      if (!(kappa_ext > tiny_dp)) continue;
      for (int psup = 1; psup <= n_theta_rt; ++psup)
        for (int k = 1; k <= n_az_rt; ++k) {
          auto I_scatt = [&](int itype) {
            const size_t q = (size_t)(k - 1) + (size_t)az_dim * ((size_t)(psup - 1) + (size_t)th_dim * ((size_t)(itype - 1) + (size_t)N_type_flux * ((size_t)(iRT - 1) + (size_t)n_RT * (size_t)(icell - 1))));
            return (double)xI[q] * factor * kappa_sca;
          };
          auto E = [&](int itype) -> double& { return eps[(size_t)(k - 1) + (size_t)az_dim * ((size_t)(psup - 1) + (size_t)th_dim * ((size_t)(itype - 1) + (size_t)N_type_flux * (size_t)(icell - 1)))]; };
          E(1) = (I_scatt(1) + J_th[icell - 1]) / kappa_ext;
          if (lsepar_pola) for (int it = 2; it <= 4; ++it) E(it) = I_scatt(it) / kappa_ext;
          if (lsepar_contrib) {
            E(n_Stokes + 2) = I_scatt(n_Stokes + 2) / kappa_ext;
            E(n_Stokes + 3) = J_th[icell - 1] / kappa_ext;
            E(n_Stokes + 4) = I_scatt(n_Stokes + 4) / kappa_ext;
          }
        }
    }
  }
  // =====================================================================
  // optical_depth.f90:1327-1421  integ_ray_dust with dust_source_fct of method 1 (dust_ray_tracing.f90:1458-1485): the
  // formal solution along a ray that is followed backwards from the observer's side
  // =====================================================================
  void integ_ray_dust(int lambda, int icell_in, double x, double y, double z, double u, double v, double w, float tau_dark_zone_obs,
                      const double* eps, int az_dim, int th_dim, int n_az_rt, int N_type_flux, double* out) {
    double x0 = x, y0 = y, z0 = z, x1 = x, y1 = y, z1 = z, l, l_contrib, l_void_before;
    int next_cell = icell_in, icell, previous_cell, icell_star = 0, i_star = 0;
    bool lintersect_stars = false;
    double tau = 0.0;
    for (int it = 0; it < N_type_flux; ++it) out[it] = 0.0;
    intersect_stars(x, y, z, u, v, w, lintersect_stars, i_star, icell_star);
    for (;;) {
      icell = next_cell;
      x0 = x1; y0 = y1; z0 = z1;
      const bool lcell_not_empty = icell <= g.n_cells;
      if (test_exit_grid(icell, x0, y0, z0)) return;
      if (lintersect_stars && icell == icell_star) return;
      previous_cell = 0;
      cross_cell(x0, y0, z0, u, v, w, icell, previous_cell, x1, y1, z1, next_cell, l, l_contrib, l_void_before);
      if (lcell_not_empty) {
        const int p_icell = lvariable_dust() ? icell : 1;
        const double dtau = l_contrib * kappa(p_icell, lambda) * kappa_factor(icell);
        const double xm = 0.5 * (x0 + x1), ym = 0.5 * (y0 + y1), zm = 0.5 * (z0 + z1);
        int k = 1, psup = 1;
        if (!g.l3D) {
          psup = (zm > 0.0) ? 1 : 2;
          const double phi_pos = std::atan2(xm, ym);
          k = (int)std::floor(fmodulo(phi_pos, two_pi) / two_pi * n_az_rt) + 1;
          if (k > n_az_rt) k = n_az_rt;
        }
        const double wgt = std::exp(-tau) * (1.0 - std::exp(-dtau));
        for (int it = 1; it <= N_type_flux; ++it)
          out[it - 1] = out[it - 1] + wgt * eps[(size_t)(k - 1) + (size_t)az_dim * ((size_t)(psup - 1) + (size_t)th_dim * ((size_t)(it - 1) + (size_t)N_type_flux * (size_t)(icell - 1)))];
        tau = tau + dtau;
        if (tau > tau_dark_zone_obs) return;
      }
    }
  }

  // =====================================================================
  // thermal_emission.f90:1771-1949  repartition_energie(lambda), LTE case: thermal emission of every cell at wavelength
  // lambda from Tdust (`real`), the cumulative emission probability of the cells and the star / disk / ISM fractions.
  // prob: prob_E_cell(0:n_cells) of this wavelength.  weight: weight_proba_emission or null; *weight_norm receives
  // prob_E_cell(n_cells) / E_disk before the normalisation (the factor correct_E_emission is multiplied by, :1932-1934).
  // Returns false when there is no energy at this wavelength (the reference exits, :1900-1904).
  // =====================================================================
  bool repartition_energie(int lambda, const float* Tdust, const double* tab_lambda, double E_star, double E_ISM, const double* weight,
                           double* prob, double& E_disk, double& frac_E_stars, double& frac_E_disk, double* weight_norm) const {
    const float thermal_const = (float)(299792458.0 * 6.626070040e-34 / 1.38064852e-23);
    const double cst_wl_max = (double)(std::log(std::numeric_limits<float>::max()) - 1.0e-4f);
    const double wl = tab_lambda[lambda - 1] * (double)1.e-6f;
    const double wl2 = wl * wl, wl5 = (wl2 * wl2) * wl;
    const int n = g.n_cells;
    E_disk = 0.0;
    prob[0] = 0.0;
    for (int icell = 1; icell <= n; ++icell) {
      double E = 0.0;
      if (!dark[icell]) {
        const double Temp = (double)Tdust[icell - 1];
        if (!(Temp < tiny_real)) {
          const double cst_wl = (double)thermal_const / (Temp * wl);
          if (cst_wl < cst_wl_max) {
            const int p_icell = lvariable_dust() ? icell : 1;
            E = 4.0 * kappa_abs_LTE(p_icell, lambda) * kappa_factor(icell) * volume(icell) / (wl5 * (std::exp(cst_wl) - 1.0));
          }
        }
      }
      E_disk = E_disk + E;
      prob[icell] = prob[icell - 1] + (weight ? E * weight[icell - 1] : E);
    }
    if (E_star + E_disk + E_ISM < tiny_dp) return false;
    frac_E_stars = E_star / (E_star + E_disk + E_ISM);
    frac_E_disk = (E_star + E_disk) / (E_star + E_disk + E_ISM);
    if (weight_norm) *weight_norm = E_disk > 0.0 ? prob[n] / E_disk : 0.0;
    const double tot = prob[n];
    if (tot > tiny_dp) for (int icell = 0; icell <= n; ++icell) prob[icell] = prob[icell] / tot;
    else for (int icell = 0; icell <= n; ++icell) prob[icell] = 0.0;
    return true;
  }

  // =====================================================================
  // scattering.f90:1354-1383  hg
  // =====================================================================
  static void hg(float g_, float rand, int& itheta, double& cospsi) {
    double rand_dp = std::min((double)rand, 1.0 - 1e-6);
    if (std::fabs(g_) > tiny_real) {
      double g1 = g_, g2 = g1 * g1;
      double q = (1.0 - g2) / (1.0 - g1 + 2.0 * g1 * rand_dp);
      cospsi = (1.0 + g2 - q * q) / (2.0 * g1);
    } else cospsi = 2.0 * rand_dp - 1.0;
    itheta = (int)std::floor(std::acos(cospsi) * 180.0 / pi) + 1;
    if (itheta > nang_scatt) itheta = nang_scatt;
  }
  // =====================================================================
  // scattering.f90:1433-1475  angle_diff_theta_pos
  // =====================================================================
  void angle_diff_theta_pos(int lambda, int icell, float rand, float rand2, int& itheta, double& cospsi) const {
    int kmin = 0, kmax = nang_scatt, k = (kmin + kmax) / 2;
    while ((kmax - kmin) > 1) {
      if (o.prob_s11_pos[pos_idx(k, icell, lambda)] < rand) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
    }
    k = kmax;
    itheta = k;
    cospsi = std::cos(((double)k - 1.0) * pi / (double)nang_scatt) +
             rand2 * (std::cos(((double)k) * pi / (double)nang_scatt) - std::cos(((double)k - 1.0) * pi / (double)nang_scatt));
  }
  // =====================================================================
  // scattering.f90:1328-1350  get_Mueller_matrix_per_cell
  // =====================================================================
  void get_Mueller_matrix_per_cell(int lambda, int itheta, float frac, int icell, double M[4][4]) const {
    float frac_m1 = 1.0f - frac;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) M[a][b] = 0.0;
    size_t q1 = pos_idx(itheta, icell, lambda), q0 = pos_idx(itheta - 1, icell, lambda);
    M[0][0] = (double)1.0f;
    M[1][1] = o.tab_s22_o_s11_pos[q1] * frac + o.tab_s22_o_s11_pos[q0] * frac_m1;     // fp32 expression
    M[0][1] = o.tab_s12_o_s11_pos[q1] * frac + o.tab_s12_o_s11_pos[q0] * frac_m1;
    M[1][0] = M[0][1];
    M[2][2] = o.tab_s33_o_s11_pos[q1] * frac + o.tab_s33_o_s11_pos[q0] * frac_m1;
    M[3][3] = o.tab_s44_o_s11_pos[q1] * frac + o.tab_s44_o_s11_pos[q0] * frac_m1;
    M[2][3] = -o.tab_s34_o_s11_pos[q1] * frac - o.tab_s34_o_s11_pos[q0] * frac_m1;
    M[3][2] = -M[2][3];
  }
  // =====================================================================
  // scattering.f90:1187-1298  update_Stokes  (fp32 angles)
  // =====================================================================
  static void update_Stokes(double* S, double u0, double v0, double w0, double u1, double v1, double w1, const double M[4][4]) {
    float sinw, cosw, omega, theta, costhet, xnyp;
    double v1pi, v1pj, v1pk, S1_0;
    rotation(u0, v0, w0, u1, v1, w1, v1pi, v1pj, v1pk);
    xnyp = (float)std::sqrt(v1pk * v1pk + v1pj * v1pj);
    if (xnyp < 1e-10f) { xnyp = 0.0f; costhet = 1.0f; }
    else costhet = (float)(-1.0f * v1pj / (double)xnyp);      // -1.0*v1pj / xnyp : dp expression stored to real
    theta = std::acos(costhet);
    if ((double)theta >= pi) theta = 0.0f;
    theta = (float)((double)theta + half_pi);
    omega = 2.0f * theta;
    if (v1pk < 0.0) omega = -1.0f * omega;
    cosw = std::cos(omega); sinw = std::sin(omega);
    if (std::fabs(cosw) < 1e-06f) cosw = 0.0f;
    if (std::fabs(sinw) < 1e-06f) sinw = 0.0f;
    double RPO[4][4] = {{0}}, ROP[4][4] = {{0}};
    RPO[0][0] = 1.0; ROP[0][0] = 1.0;
    RPO[1][1] = cosw; ROP[1][1] = cosw;
    RPO[1][2] = sinw; ROP[2][1] = sinw;
    RPO[2][1] = -1.0f * sinw; ROP[1][2] = -1.0f * sinw;
    RPO[2][2] = cosw; ROP[2][2] = cosw;
    RPO[3][3] = 1.0; ROP[3][3] = 1.0;
    S1_0 = S[0];
    double C[4], D[4], R[4];
    for (int a = 0; a < 4; ++a) { C[a] = 0; for (int b = 0; b < 4; ++b) C[a] += ROP[a][b] * S[b]; }
    for (int a = 0; a < 4; ++a) { D[a] = 0; for (int b = 0; b < 4; ++b) D[a] += M[a][b] * C[b]; }
    for (int a = 0; a < 4; ++a) { R[a] = 0; for (int b = 0; b < 4; ++b) R[a] += RPO[a][b] * D[b]; }
    for (int a = 0; a < 4; ++a) S[a] = R[a];
    if (S[0] > tiny_real) { const double S0 = S[0]; for (int a = 0; a < 4; ++a) S[a] = S[a] * M[0][0] * S1_0 / S0; }   // :1294 left-to-right
  }

  // =====================================================================
  // thermal_emission.f90:649-706  Temp_LTE   (id > 0 branch; id == 0 in temp_finale)
  // `frac` is intent(out) but left undefined by the reference when the cell is
  // at T_min (:673-679); the oracle returns frac = 0 there (documented choice).
  // =====================================================================
  void Temp_LTE(ThreadTallies& t, int icell, int& Ti, float& Temp, double& frac) {
    int p_icell = lvariable_dust() ? icell : 1;
    double Qheat = t.xKJ_abs[icell - 1] * nb_proc * e.L_packet_th / volume(icell);
    frac = 0.0;
    if (Qheat < tiny_dp) { Temp = o.T_min; Ti = 2; }
    else {
      double log_Qheat = std::log(Qheat);
      if (log_Qheat < log_Qcool(1, p_icell)) { Temp = o.T_min; Ti = 2; }
      else {
        Ti = t.xT_ech[icell - 1];
        while ((log_Qcool(Ti, p_icell) < log_Qheat) && (Ti < o.n_T)) Ti = Ti + 1;
        frac = (log_Qheat - log_Qcool(Ti - 1, p_icell)) / (log_Qcool(Ti, p_icell) - log_Qcool(Ti - 1, p_icell));
        Temp = (float)std::exp((double)std::log(tab_Temp(Ti)) * frac + (double)std::log(tab_Temp(Ti - 1)) * (1.0 - frac));
      }
    }
    t.xT_ech[icell - 1] = Ti;
  }
  // =====================================================================
  // thermal_emission.f90:710-771  im_reemission_LTE  (high-memory branch)
  // =====================================================================
  void im_reemission_LTE(ThreadTallies& t, int icell, int p_icell, float rand1, float rand2, int& lambda) {
    int Ti; float Temp; double frac_T2;
    Temp_LTE(t, icell, Ti, Temp, frac_T2);
    int T2 = Ti, T1 = Ti - 1;
    double frac_T1 = 1.0 - frac_T2;
    int l1 = 0, l2 = o.n_lambda, l = (l1 + l2) / 2;
    if (r.low_mem_th_emission) {                                       // :739-751
      const int k = select_absorbing_grain(lambda, icell, rand1, 1);
      const int nk = gr.grain_RE_LTE_end - gr.grain_RE_LTE_start + 1;
      auto cdf = [&](int ll, int Tt) { return gr.kdB_dT_1grain_LTE_CDF[(size_t)(ll - 1) + (size_t)o.n_lambda * ((size_t)(k - gr.grain_RE_LTE_start) + (size_t)nk * (Tt - 1))]; };
      while ((l2 - l1) > 1) {
        double proba = frac_T1 * cdf(l, T1) + frac_T2 * cdf(l, T2);
        if ((double)rand2 > proba) l1 = l; else l2 = l;
        l = (l1 + l2) / 2;
      }
      lambda = l + 1;
      return;
    }
    while ((l2 - l1) > 1) {
      double proba = frac_T1 * kdB_dT_CDF(l, T1, p_icell) + frac_T2 * kdB_dT_CDF(l, T2, p_icell);
      if ((double)rand2 > proba) l1 = l; else l2 = l;
      l = (l1 + l2) / 2;
    }
    lambda = l + 1;
  }


  // =====================================================================
  // thermal_emission.f90:1953-2040  select_absorbing_grain  (heating_method 2, 3)
  // =====================================================================
  int select_absorbing_grain(int lambda, int icell, float rand, int heating_method) const {
    const double AU3 = AU_to_cm_mum2;
    int p_icell = lvariable_dust() ? icell : 1;          // icell1
    double norm; int kstart, kend, k;
    if (heating_method == 1) {
      norm = kappa_abs_LTE(p_icell, lambda) * kappa_factor(icell) / AU3;
      kstart = gr.grain_RE_LTE_start; kend = gr.grain_RE_LTE_end;
    } else if (heating_method == 2) {
      norm = kappa_abs_nLTE(p_icell, lambda) * kappa_factor(icell) / AU3;
      kstart = gr.grain_RE_nLTE_start; kend = gr.grain_RE_nLTE_end;
    } else {
      if (r.lRE_nLTE) norm = (gr.kappa_abs_RE[cl_idx(icell, lambda)] - (kappa_abs_LTE(p_icell, lambda) + kappa_abs_nLTE(p_icell, lambda)) * kappa_factor(icell)) / AU3;
      else            norm = (gr.kappa_abs_RE[cl_idx(icell, lambda)] - kappa_abs_LTE(p_icell, lambda) * kappa_factor(icell)) / AU3;
      kstart = gr.grain_nRE_start; kend = gr.grain_nRE_end;
    }
    // p_k => k if lvariable_dust else => k1 = 1  (:1970-1977: zone 1 whatever the grain)
    // C_abs(k,lambda) * dust_density_o_n_grains(p_k,icell) * n_grains(k), left to right
    auto term = [&](int kk) { return (double)gr.C_abs[gl_idx(kk, lambda)] * dust_density_o_n_grains(lvariable_dust() ? kk : 1, icell) * gr.n_grains[kk - 1]; };
    const bool masked = heating_method > 2;
    double prob, CDF = 0.0;
    if (rand < 0.5f) {
      prob = rand * norm;
      for (k = kstart; k <= kend; ++k) {
        if (!masked || l_RE(k, icell)) CDF = CDF + term(k);
        if (CDF > prob) break;
      }
    } else {
      prob = (1.0f - rand) * norm;                      // (1.0-rand) evaluated in fp32
      for (k = kend; k >= kstart; --k) {
        if (!masked || l_RE(k, icell)) CDF = CDF + term(k);
        if (CDF > prob) break;
      }
    }
    return k;       // kend+1 / kstart-1 when the loop runs out, exactly like the Fortran do-variable
  }

  // shared tail of im_reemission_NLTE / im_reemission_qRE: temperature of grain k from
  // log_E_abs, then the wavelength bisection  (thermal_emission.f90:812-863, 1469-1511)
  template <class FE, class FC>
  void reemit_1grain(std::vector<int> ThreadTallies::*xT, ThreadTallies& t, size_t ix, double log_E_abs, float rand2, FE logE, FC cdf, int& lambda) {
    int T_int = (T[0].*xT)[ix];
    for (auto& tt : T) T_int = std::max(T_int, (tt.*xT)[ix]);            // maxval(xT_ech_1grain(k,icell,:))
    while ((logE(T_int) < log_E_abs) && (T_int < o.n_T)) T_int = T_int + 1;
    (t.*xT)[ix] = T_int;
    int T2 = T_int, T1 = T_int - 1;
    double Temp2 = tab_Temp(T2), Temp1 = tab_Temp(T1);
    double frac = (log_E_abs - logE(T1)) / (logE(T2) - logE(T1));
    double Temp = std::exp(std::log(Temp2) * frac + std::log(Temp1) * (1.0 - frac));
    double frac_T2 = (Temp - Temp1) / (Temp2 - Temp1), frac_T1 = 1.0 - frac_T2;
    int l1 = 0, l2 = o.n_lambda, l = (l1 + l2) / 2;
    while ((l2 - l1) > 1) {
      double proba = frac_T1 * cdf(l, T1) + frac_T2 * cdf(l, T2);
      if ((double)rand2 > proba) l1 = l; else l2 = l;
      l = (l1 + l2) / 2;
    }
    lambda = l + 1;
  }
  // sum(xJ_abs(icell,ilambda,:)) over the id slices
  double xJ_abs_sum(int icell, int ilambda) const { double s = 0.0; size_t ix = cl_idx(icell, ilambda); for (auto& tt : T) s += tt.xJ_abs[ix]; return s; }

  // =====================================================================
  // thermal_emission.f90:775-866  im_reemission_NLTE
  // =====================================================================
  void im_reemission_NLTE(ThreadTallies& t, int icell, int /*p_icell*/, float rand1, float rand2, int& lambda) {
    const int lambda0 = lambda;
    int k;
    if (r.low_mem_th_emission_nLTE) k = select_absorbing_grain(lambda0, icell, rand1, 2);
    else {
      int kmin = gr.grain_RE_nLTE_start, kmax = gr.grain_RE_nLTE_end;
      k = (kmin + kmax) / 2;
      while ((kmax - kmin) > 1) {
        if (kabs_nLTE_CDF(k, icell, lambda0) < (double)rand1) kmin = k; else kmax = k;
        k = (kmin + kmax) / 2;
      }
      k = kmax;
    }
    double J_abs = 0.0;
    for (int il = 1; il <= o.n_lambda; ++il) J_abs = J_abs + gr.C_abs_norm[gl_idx(k, il)] * (xJ_abs_sum(icell, il) + gr.J0[cl_idx(icell, il)]);
    double log_E_abs = std::log(J_abs * e.L_packet_th / volume(icell));
    size_t ix = (size_t)(k - gr.grain_RE_nLTE_start) + (size_t)nk_nLTE() * (icell - 1);
    reemit_1grain(&ThreadTallies::xT_ech_1grain, t, ix, log_E_abs, rand2,
                  [&](int Ti) { return log_E_em_1grain(k, Ti); }, [&](int l, int Ti) { return kdB_dT_1grain_nLTE_CDF(l, k, Ti); }, lambda);
  }
  // =====================================================================
  // thermal_emission.f90:1441-1514  im_reemission_qRE
  // (J0(icell,lambda) is indexed with the absorbed wavelength, not ilambda: :1466)
  // =====================================================================
  void im_reemission_qRE(ThreadTallies& t, int icell, int /*p_icell*/, float rand1, float rand2, int& lambda) {
    const int lambda0 = lambda;
    int k = select_absorbing_grain(lambda0, icell, rand1, 3);
    double J_abs = 0.0;
    for (int il = 1; il <= o.n_lambda; ++il) J_abs = J_abs + gr.C_abs_norm[gl_idx(k, il)] * (xJ_abs_sum(icell, il) + gr.J0[cl_idx(icell, lambda)]);
    double log_E_abs = std::log(J_abs * e.L_packet_th / volume(icell));
    size_t ix = (size_t)(k - gr.grain_nRE_start) + (size_t)nk_nRE() * (icell - 1);
    reemit_1grain(&ThreadTallies::xT_ech_1grain_nRE, t, ix, log_E_abs, rand2,
                  [&](int Ti) { return log_E_em_1grain_nRE(k, Ti); }, [&](int l, int Ti) { return kdB_dT_1grain_nRE_CDF(l, k, Ti); }, lambda);
  }

  // =====================================================================
  // dust_prop.f90:1292-1336 select_scattering_grain, :1340-1380 select_grainsize_high_mem
  // =====================================================================
  int select_scattering_grain(int lambda, int icell, float rand) const {
    if (r.low_mem_scattering) {
      double norm = kappa(icell, lambda) * tab_albedo_pos(icell, lambda) / AU_to_cm_mum2;
      double prob, CDF = 0.0; int k;
      auto dens = [&](int kk) { return dust_density_o_n_grains(lvariable_dust() ? kk : gr.grain_zone[kk - 1], icell) * gr.n_grains[kk - 1]; };
      if (rand < 0.5f) {
        prob = rand * norm;
        for (k = 1; k <= gr.n_grains_tot; ++k) { CDF = CDF + gr.C_sca[gl_idx(k, lambda)] * dens(k); if (CDF > prob) break; }
      } else {
        prob = (1.0f - rand) * norm;
        for (k = gr.n_grains_tot; k >= 1; --k) { CDF = CDF + gr.C_sca[gl_idx(k, lambda)] * dens(k); if (CDF > prob) break; }
      }
      return k;
    }
    float prob = rand;
    int kmin = 0, kmax = gr.n_grains_tot, k = (kmin + kmax) / 2;
    while (ksca_CDF(k, icell, lambda) != (double)prob) {
      if (ksca_CDF(k, icell, lambda) < (double)prob) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
      if ((kmax - kmin) <= 1) break;
    }
    return kmax;
  }
  // =====================================================================
  // scattering.f90:1387-1429  angle_diff_theta  (per grain)
  // =====================================================================
  void angle_diff_theta(int lambda, int igrain, float rand, float rand2, int& itheta, double& cospsi) const {
    int kmin = 0, kmax = nang_scatt, k = (kmin + kmax) / 2;
    while ((kmax - kmin) > 1) {
      if (prob_s11(lambda, igrain, k) < rand) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
    }
    k = kmax;
    itheta = k;
    // real(k), real(nang_scatt) are fp32 but promoted: (real(k)-1.0)*pi/real(nang_scatt) is dp because pi is dp
    cospsi = std::cos(((double)((float)k - 1.0f)) * pi / (double)nang_scatt) +
             rand2 * (std::cos(((double)(float)k) * pi / (double)nang_scatt) - std::cos(((double)((float)k - 1.0f)) * pi / (double)nang_scatt));
  }
  // =====================================================================
  // scattering.f90:1302-1324  get_Mueller_matrix_per_grain
  // =====================================================================
  void get_Mueller_matrix_per_grain(int lambda, int itheta, float frac, int igrain, double M[4][4]) const {
    float frac_m1 = 1.0f - frac;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) M[a][b] = 0.0;
    size_t q1 = s_idx(itheta, igrain, lambda), q0 = s_idx(itheta - 1, igrain, lambda);
    M[0][0] = gr.tab_s11[q1] * frac + gr.tab_s11[q0] * frac_m1;          // fp32 expressions
    M[1][1] = gr.tab_s22[q1] * frac + gr.tab_s22[q0] * frac_m1;
    M[0][1] = gr.tab_s12[q1] * frac + gr.tab_s12[q0] * frac_m1;
    M[1][0] = M[0][1];
    M[2][2] = gr.tab_s33[q1] * frac + gr.tab_s33[q0] * frac_m1;
    M[3][3] = gr.tab_s44[q1] * frac + gr.tab_s44[q0] * frac_m1;
    M[2][3] = -gr.tab_s34[q1] * frac - gr.tab_s34[q0] * frac_m1;
    M[3][2] = -M[2][3];
  }

  // =====================================================================
  // dust_transfer.f90:1155-1409  propagate_packet  
  // =====================================================================
  void propagate_packet(PacketRng& rng, ThreadTallies& t, int& lambda, int p_lambda, Packet& p) {
    double M[4][4], u1, v1, w1, phi, cospsi;
    int itheta;
    float rand, rand2, tau, dvol;
    bool flag_direct_star, flag_sortie = false;
    p.flag_scatt = false;
    flag_direct_star = p.flag_star;
    uint32_t n_flight = 0;
    int n_iteractions_in_cell = 0;                // :1204
    for (;;) {
      ++n_flight;
      if (r.lMRW && n_iteractions_in_cell > 5) {  // :1222-1239
        if (modified_random_walk(rng, t, n_flight, lambda, p)) flag_direct_star = false;
      }
      rng_set_block(&rng, 2 * n_flight);          // flight block: tau, interaction-type draw (see philox.h)
      rand = (float)rng_next(&rng);
      if (rand == 1.0f) tau = 1.0e30f;
      else if (rand > 1.0e-6f) tau = -std::log(1.0f - rand);
      else tau = rand;
      const int icell_old = p.icell;                                    // :1242-1249
      physical_length(t, lambda, p_lambda, p.S, p.icell, p.x, p.y, p.z, p.u, p.v, p.w, p.flag_star, flag_direct_star, tau, dvol, flag_sortie, p.alive);
      if (p.icell == icell_old) n_iteractions_in_cell = n_iteractions_in_cell + 1;
      else n_iteractions_in_cell = 0;
      if (flag_sortie) return;
      int p_icell = lvariable_dust() ? p.icell : 1;
      t.stats[2] += 1;
      flag_direct_star = false;
      if (r.lmono) {
        if (dark[p.icell]) { p.alive = false; return; }
        float alb = tab_albedo_pos(p_icell, lambda);
        for (int a = 0; a < 4; ++a) p.S[a] = p.S[a] * alb;
        if (p.S[0] < tiny_real_x1e6) { p.alive = false; return; }
        rand = -1.0f;
      } else rand = (float)rng_next(&rng);

      rng_set_block(&rng, 2 * n_flight + 1);      // interaction block
      if (rand < tab_albedo_pos(p_icell, lambda)) {
        t.stats[3] += 1; logev(2);
        p.flag_scatt = true; flag_direct_star = false;
        if (r.lscattering_method1) {                 // :1291-1317
          rand = (float)rng_next(&rng);
          int igrain = select_scattering_grain(lambda, p_icell, rand);
          rand = (float)rng_next(&rng);
          rand2 = (float)rng_next(&rng);
          if (r.lmethod_aniso1) {
            angle_diff_theta(lambda, igrain, rand, rand2, itheta, cospsi);
            rand = (float)rng_next(&rng);
            phi = pi * (double)(2.0f * rand - 1.0f);
            cdapres(cospsi, phi, p.u, p.v, p.w, u1, v1, w1);
            if (r.lsepar_pola) {
              get_Mueller_matrix_per_grain(lambda, itheta, rand2, igrain, M);
              update_Stokes(p.S, p.u, p.v, p.w, u1, v1, w1, M);
            }
          } else {
            hg(gr.tab_g[gl_idx(igrain, lambda)], rand, itheta, cospsi);
            if (r.lisotropic) { itheta = 1; cospsi = (double)(2.0f * rand - 1.0f); }
            rand = (float)rng_next(&rng);
            phi = pi * (double)(2.0f * rand - 1.0f);
            cdapres(cospsi, phi, p.u, p.v, p.w, u1, v1, w1);
          }
          p.u = u1; p.v = v1; p.w = w1;
          continue;
        }
        // method 2  :1318-1348
        rand = (float)rng_next(&rng);
        rand2 = (float)rng_next(&rng);
        if (r.lmethod_aniso1) {
          angle_diff_theta_pos(p_lambda, p_icell, rand, rand2, itheta, cospsi);
          if (r.lisotropic) { itheta = 1; cospsi = (double)(2.0f * rand - 1.0f); }    // :1325 fp32 expression
          rand = (float)rng_next(&rng);
          phi = pi * (double)(2.0f * rand - 1.0f);                                      // :1329 PI*(2.0*rand-1.0): fp32 inner
          cdapres(cospsi, phi, p.u, p.v, p.w, u1, v1, w1);
          if (r.lsepar_pola) {
            get_Mueller_matrix_per_cell(lambda, itheta, rand2, p_icell, M);
            update_Stokes(p.S, p.u, p.v, p.w, u1, v1, w1, M);
          }
        } else {
          hg(tab_g_pos(p_icell, lambda), rand, itheta, cospsi);
          if (r.lisotropic) { itheta = 1; cospsi = (double)(2.0f * rand - 1.0f); }    // :1340
          rand = (float)rng_next(&rng);
          phi = pi * (double)(2.0f * rand - 1.0f);                                      // :1344
          cdapres(cospsi, phi, p.u, p.v, p.w, u1, v1, w1);
        }
        p.u = u1; p.v = v1; p.w = w1;
      } else {
        t.stats[4] += 1; logev(3);
        if ((!r.lmono) && r.lnRE) {                                     // :1355-1366
          const double pRE = gr.proba_abs_RE[cl_idx(p.icell, lambda)];
          t.E_abs_nRE = t.E_abs_nRE + p.S[0] * (1.0 - pRE);
          for (int a = 0; a < 4; ++a) p.S[a] = p.S[a] * pRE;
          if (p.S[0] < tiny_real) { p.alive = false; return; }
        }
        p.flag_star = false; p.flag_scatt = false; flag_direct_star = false; p.flag_ISM = false;
        if (r.lonly_LTE) {
          rand = (float)rng_next(&rng); rand2 = (float)rng_next(&rng);   // :1373-1375
          im_reemission_LTE(t, p.icell, p_icell, rand, rand2, lambda);
        } else if (r.lonly_nLTE) {
          rand = (float)rng_next(&rng); rand2 = (float)rng_next(&rng);
          im_reemission_NLTE(t, p.icell, p_icell, rand, rand2, lambda);
        } else {
          // grain-regime draw: its own Philox block so that rand/rand2/direction keep their words
          rng_set_block(&rng, (2 * n_flight + 1) | 0x80000000u);
          rand = (float)rng_next(&rng);
          rng_set_block(&rng, 2 * n_flight + 1);
          const float sel = rand;
          rand = (float)rng_next(&rng); rand2 = (float)rng_next(&rng);
          if ((double)sel <= gr.Proba_abs_RE_LTE[cl_idx(p.icell, lambda)]) im_reemission_LTE(t, p.icell, p_icell, rand, rand2, lambda);
          else if ((double)sel <= gr.Proba_abs_RE_LTE_p_nLTE[cl_idx(p.icell, lambda)]) im_reemission_NLTE(t, p.icell, p_icell, rand, rand2, lambda);
          else im_reemission_qRE(t, p.icell, p_icell, rand, rand2, lambda);
        }
        random_isotropic_direction(rng, p.u, p.v, p.w);
        p.S[1] = 0.0; p.S[2] = 0.0; p.S[3] = 0.0;
      }
    }
  }


  // =====================================================================
  // Modified random walk.  MRW.f90:16-115 (zeta table, sample_zeta, make_MRW_step -- unfinished in the
  // reference), call site dust_transfer.f90:1222-1239 (commented out), distance_to_closest_wall_*
  // (cylindrical_grid.f90:1179-1226, spherical_grid.f90:451-499, Voronoi.f90:996-1061), mean opacities
  // compute_Planck_opacities diffusion.f90:631-693 (T hard-wired to 20 K there).  The step itself follows
  // Min et al. 2009 (A&A 497, 155) sect. 3-4 and Robitaille 2010 (A&A 520, A70), with the mean opacities
  // taken over the spectrum this code's own immediate re-emission samples (kdB_dT_CDF at the cell's running
  // temperature), so that the walk it replaces is exactly the one propagate_packet would have done.
  // =====================================================================
  static constexpr int n_zeta = 10000;            // MRW.f90:8
  std::vector<double> zeta_tab;                    // zeta(1:n), y_MRW(i) = (i-1)/(n-1)
  std::vector<double> mrw_A, mrw_B, mrw_C;         // (n_T, p_n_cells)
  const void* mrw_for = nullptr;                   // opacity tables the means were built from

  // MRW.f90:16-54 initialize_cumulative_zeta: zeta(y) = 2 Sum_{n>=1} (-1)^(n+1) y^(n^2)  (Min et al. 2009 eq. 7).
  // A running maximum is applied afterwards: beyond y ~ 0.95 the series equals 1 to rounding and its noise would
  // make the table non-monotonic (no 24-bit draw can land there; it keeps the lookup well defined).
  void initialize_cumulative_zeta() {
    zeta_tab.assign(n_zeta, 0.0);
    for (int i = 1; i <= n_zeta; ++i) {
      const double y = (double)(i - 1) / (double)(n_zeta - 1);
      double z = 0.0;
      if (i == n_zeta) z = 0.5;
      else {
        int j = 0;
        for (;;) {
          j = j + 1;
          const double term = std::pow(y, (double)j * (double)j);
          if (term == 0.0) break;
          if (j % 2 == 0) z = z - term; else z = z + term;
        }
      }
      zeta_tab[i - 1] = z * 2.0;
    }
    for (int i = 1; i < n_zeta; ++i) zeta_tab[i] = std::max(zeta_tab[i], zeta_tab[i - 1]);
  }
  // MRW.f90:58-70 sample_zeta = interp(y_MRW, zeta, zeta_random)  (utils.f90:190-247: first j in 2..n-1 with
  // x(j) > xp, else n; linear interpolation, no extrapolation)
  double sample_zeta(double zeta_random) const {
    const int n = n_zeta;
    if (zeta_random < zeta_tab[0]) return 0.0;
    if (zeta_random > zeta_tab[n - 1]) return 1.0;
    int j = 2;
    for (; j <= n - 1; ++j) if (zeta_tab[j - 1] > zeta_random) break;
    const double x0 = zeta_tab[j - 2], x1 = zeta_tab[j - 1];
    const double y0 = (double)(j - 2) / (double)(n - 1), y1 = (double)(j - 1) / (double)(n - 1);
    const double frac = (zeta_random - x0) / (x1 - x0);
    return y0 * (1. - frac) + y1 * frac;
  }

  // sin_phi_lim(k) as a wall normal: the grid set-up stores 1.0d300 next to tan_phi_lim = 1.0d300 for the walls at
  // phi = pi/2 (mod pi) (cylindrical_grid.f90:591-594), which would take those walls out of the minimum; they are the
  // planes x = 0, i.e. sin = 1, cos = 0.
  inline double sin_phi_wall(int k) const { const double sp = g.sin_phi_lim[k - 1]; return sp > 1.0e299 ? 1.0 : sp; }
  // cylindrical_grid.f90:1179-1226 distance_to_closest_wall_cyl.  (k0-1 = 0 indexes sin_phi_lim(0) in the
  // reference, out of bounds; the wall below sector 1 is the upper wall of sector n_az.)
  double distance_to_closest_wall_cyl(int icell, double x, double y, double z) const {
    const int ri0 = cmap_i[icell]; int zj0 = cmap_j[icell]; const int k0 = cmap_k[icell];
    const double rr = std::sqrt(x * x + y * y);
    const double s1 = r_lim(ri0) - rr;
    const double s2 = rr - r_lim(ri0 - 1);
    const double z0 = std::fabs(z);
    zj0 = std::abs(zj0);
    const double s3 = z_lim(ri0, std::abs(zj0) + 1) - z0;
    const double s4 = z0 - z_lim(ri0, std::abs(zj0));
    double s_ = std::min(std::min(s1, s2), std::min(s3, s4));
    if (g.l3D) {
      const int km = (k0 - 1 >= 1) ? k0 - 1 : g.n_az;
      const double s5 = std::fabs(x * sin_phi_wall(k0) - y * g.cos_phi_lim[k0 - 1]);
      const double s6 = std::fabs(x * sin_phi_wall(km) - y * g.cos_phi_lim[km - 1]);
      s_ = std::min(s_, std::min(s5, s6));
    }
    return s_;
  }
  // spherical_grid.f90:451-499 distance_to_closest_wall_sph.  The reference writes the theta walls as
  // abs(rcyl*w_lim(thetaj0) - z0*cos_phi_lim(thetaj0)) (:471-472): cos_phi_lim is the AZIMUTHAL table (size n_az,
  // zero in 2D) indexed with the theta index -- an out-of-bounds read of dead code.  The distance from (rcyl, z0)
  // to the cone z = rcyl tan(theta_lim(j)) is |rcyl sin(theta_lim) - z0 cos(theta_lim)|; that is what is computed,
  // with cos(theta_lim(j)) in place of the mis-indexed table.
  double distance_to_closest_wall_sph(int icell, double x, double y, double z) const {
    const int ri0 = cmap_i[icell]; const int thetaj0 = std::abs(cmap_j[icell]); const int k0 = cmap_k[icell];
    const double r2_cyl = x * x + y * y;
    const double rcyl = std::sqrt(r2_cyl);
    const double rr = std::sqrt(r2_cyl + z * z);
    const double s1 = r_lim(ri0) - rr;
    const double s2 = rr - r_lim(ri0 - 1);
    const double z0 = std::fabs(z);
    const double s3 = std::fabs(rcyl * g.w_lim[thetaj0] - z0 * std::cos(theta_lim(thetaj0)));
    const double s4 = std::fabs(rcyl * g.w_lim[thetaj0 - 1] - z0 * std::cos(theta_lim(thetaj0 - 1)));
    double s_ = std::min(std::min(s1, s2), std::min(s3, s4));
    if (g.l3D) {
      const int km = (k0 - 1 >= 1) ? k0 - 1 : g.n_az;
      const double s5 = std::fabs(x * sin_phi_wall(k0) - y * g.cos_phi_lim[k0 - 1]);
      const double s6 = std::fabs(x * sin_phi_wall(km) - y * g.cos_phi_lim[km - 1]);
      s_ = std::min(s_, std::min(s5, s6));
    }
    return s_;
  }
  // Voronoi.f90:996-1061 distance_to_closest_wall_Voronoi (fp32 plane geometry like cross_Voronoi_cell).
  // The reference divides dot(n, p - r) by den = dot(n, n) with the UN-normalised normal n = r_neighbour - r_cell,
  // which yields the distance in units of |n|, not a length; the length is dot(n, p - r) / |n| and that is
  // what is returned here (den = sqrt(dot(n, n))).  Cut cells and cells touching a wall return 0 (:1009, :1051).
  double distance_to_closest_wall_Voronoi(int icell, double x, double y, double z) const {
    if (g.vor_was_cut[icell - 1] != 0) return 0.0;
    float n[3], p[3], r[3], r_cell[3], r_neighbour[3];
    r[0] = (float)x; r[1] = (float)y; r[2] = (float)z;
    double s_ = (double)1e30f;
    const double* cxyz = g.vor_xyz + 3 * (size_t)(icell - 1);
    for (int a = 0; a < 3; ++a) r_cell[a] = (float)cxyz[a];
    const int ifirst = g.vor_first[icell - 1], ilast = g.vor_last[icell - 1];
    for (int i = ifirst; i <= ilast; ++i) {
      const int id_n = g.neighbours_list[i - 1];
      double s_tmp;
      if (id_n > 0) {
        const double* nxyz = g.vor_xyz + 3 * (size_t)(id_n - 1);
        for (int a = 0; a < 3; ++a) r_neighbour[a] = (float)nxyz[a];
        for (int a = 0; a < 3; ++a) n[a] = r_neighbour[a] - r_cell[a];
        const double den = (double)std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        for (int a = 0; a < 3; ++a) p[a] = 0.5f * (r_neighbour[a] + r_cell[a]);
        const float dot = n[0] * (p[0] - r[0]) + n[1] * (p[1] - r[1]) + n[2] * (p[2] - r[2]);
        s_tmp = (double)dot / den;
        if (s_tmp < 0.) s_tmp = (double)huge_real;
      } else s_tmp = 0.0;
      if (s_tmp < s_) s_ = s_tmp;
    }
    return s_;
  }
  double distance_to_closest_wall(int icell, double x, double y, double z) const {
    if (g.kind == MCB_GRID_CYL) return distance_to_closest_wall_cyl(icell, x, y, z);
    if (g.kind == MCB_GRID_SPH) return distance_to_closest_wall_sph(icell, x, y, z);
    return distance_to_closest_wall_Voronoi(icell, x, y, z);
  }

  // Mean opacities per temperature index (the job of compute_Planck_opacities, diffusion.f90:631-693).
  // One absorb / re-emit cycle of propagate_packet: a wavelength from p_T(lambda) = increments of
  // kdB_dT_CDF(:,T), then flights of mean length 1/(kf kappa) with albedo a and asymmetry g until the
  // absorption.  Per cycle, in units of kappa_factor:
  //   A(T) = Sum p / (kappa (1-a))                      mean path
  //   B(T) = Sum p / (kappa (1-a) kappa (1-a g))        mean squared displacement / 2
  //   C(T) = Sum p kappa_abs_LTE / (kappa (1-a))        mean xKJ_abs deposit / Stokes(1)
  // so that the diffusion coefficient per unit path is B / (3 A kf) (a Rosseland-type mean of the transport
  // opacity) and the energy left per unit path is C / A (a Planck-type mean of the absorption opacity).
  void compute_MRW_means() {
    if (mrw_for == (const void*)o.kdB_dT_CDF && !mrw_A.empty()) return;
    const int nT = o.n_T, npc = o.p_n_cells, nl = o.n_lambda;
    mrw_A.assign((size_t)nT * npc, 0.0); mrw_B.assign((size_t)nT * npc, 0.0); mrw_C.assign((size_t)nT * npc, 0.0);
    for (int pc = 1; pc <= npc; ++pc)
      for (int t = 1; t <= nT; ++t) {
        double A = 0.0, B = 0.0, Cc = 0.0;
        for (int l = 1; l <= nl; ++l) {
          const double pl = kdB_dT_CDF(l, t, pc) - (l > 1 ? kdB_dT_CDF(l - 1, t, pc) : 0.0);
          const double kap = kappa(pc, l), a = (double)tab_albedo_pos(pc, l);
          const double gg = o.tab_g_pos ? (double)tab_g_pos(pc, l) : 0.0;
          const double k_abs = kap * (1.0 - a), k_tr = kap * (1.0 - a * gg);
          if (!(pl > 0.0) || !(k_abs > 0.0)) continue;
          A = A + pl / k_abs;
          B = B + pl / (k_abs * k_tr);
          Cc = Cc + pl * kappa_abs_LTE(pc, l) / k_abs;
        }
        const size_t q = (size_t)(t - 1) + (size_t)nT * (pc - 1);
        mrw_A[q] = A; mrw_B[q] = B; mrw_C[q] = Cc;
      }
    mrw_for = (const void*)o.kdB_dT_CDF;
  }

  // dust_transfer.f90:1222-1239 (restated; see the header of this section).  Returns true if at least one
  // step was made; the packet then leaves with a new wavelength, an isotropic direction and no polarisation.
  // RNG: every step and the final re-emission take one event number each, block (2 n_flight + 1) | 0x40000000.
  bool modified_random_walk(PacketRng& rng, ThreadTallies& t, uint32_t& n_flight, int& lambda, Packet& p) {
    const int icell = p.icell;
    if (icell < 1 || icell > g.n_cells) return false;
    const double kf = kappa_factor(icell);
    if (!(kf > 0.0)) return false;
    const int p_icell = lvariable_dust() ? icell : 1;
    const double gamma = (r.gamma_MRW > 0.0f) ? (double)r.gamma_MRW : 2.0;
    double d = distance_to_closest_wall(icell, p.x, p.y, p.z);
    // running temperature of the cell -> mean opacities (interpolated like the re-emission CDF)
    int Ti; float Temp; double frac_T2;
    Temp_LTE(t, icell, Ti, Temp, frac_T2);
    const double frac_T1 = 1.0 - frac_T2;
    const size_t q1 = (size_t)(Ti - 2) + (size_t)o.n_T * (p_icell - 1), q2 = q1 + 1;
    const double A = frac_T1 * mrw_A[q1] + frac_T2 * mrw_A[q2];
    const double B = frac_T1 * mrw_B[q1] + frac_T2 * mrw_B[q2];
    const double Cc = frac_T1 * mrw_C[q1] + frac_T2 * mrw_C[q2];
    if (!(A > 0.0) || !(B > 0.0)) return false;
    const double l_R = B / (A * kf);                 // 1 / (rho chi_R): Rosseland-type mean free path
    bool did = false;
    int n_steps = 0;
    t.stats[10] += 1;                                // diagnostics: walks attempted / refused (d <= gamma l_R)
    if (!(d > gamma * l_R)) { t.stats[11] += 1; if (getenv("ORACLE_MRW_DEBUG") && ((int)t.stats[11] % 20000) == 1) fprintf(stderr, "refused: cell %d (ri %d zj %d) d %.4g l_R %.4g Ti %d frac %.3f lambda %d kf %.4g 1/(kf kappa(lambda)) %.4g A %.4g B %.4g\n", icell, cmap_i[icell], cmap_j[icell], d, l_R, Ti, frac_T2, lambda, kf, 1.0 / (kf * kappa(p_icell, lambda)), A, B); }
    while (d > gamma * l_R && n_steps < 100000) {
      rng_set_block(&rng, (2 * n_flight + 1) | 0x40000000u);
      // MRW.f90:84-88: random point on the sphere of radius d around the packet
      double u, v, w;
      random_isotropic_direction(rng, u, v, w);
      p.x = p.x + u * d; p.y = p.y + v * d; p.z = p.z + w * d;
      // MRW.f90:92-98: y from zeta(y) = random; Min et al. 2009 eq. 8: c t = -ln(y) (d/pi)^2 / D, D = l_R / 3
      double zr = rng_next(&rng);
      if (zr <= 0.0) zr = 1.0 / 33554432.0;
      const double ym = sample_zeta(zr);
      const double ct = -std::log(ym) * (d / pi) * (d / pi) * 3.0 / l_R;
      // energy left in the cell along the walk (Lucy path-length estimator, radiation_field.f90:53)
      t.xKJ_abs[icell - 1] = t.xKJ_abs[icell - 1] + p.S[0] * ct * Cc / A;
      t.stats[9] += 1;
      did = true; ++n_steps; ++n_flight;
      d = distance_to_closest_wall(icell, p.x, p.y, p.z);
    }
    if (!did) return false;
    t.stats[8] += 1;
    // end of the walk: thermal re-emission at the cell's running temperature
    rng_set_block(&rng, (2 * n_flight + 1) | 0x40000000u);
    const float rand1 = (float)rng_next(&rng), rand2 = (float)rng_next(&rng);
    im_reemission_LTE(t, icell, p_icell, rand1, rand2, lambda);
    random_isotropic_direction(rng, p.u, p.v, p.w);
    p.S[1] = 0.0; p.S[2] = 0.0; p.S[3] = 0.0;
    p.flag_star = false; p.flag_scatt = false; p.flag_ISM = false;
    ++n_flight;
    return true;
  }

  // =====================================================================
  // output.f90:294-595  capteur  (SED branch; lorigine / image maps not built)
  // returns capt (0 if the packet is dropped by the symmetry test)
  // =====================================================================
  int capteur(ThreadTallies& t, int lambda, const Packet& p) {
    // output.f90:308-319
    const int maxigrid = std::max(r.npix_x, r.npix_y);
    int deltapix_x = 1, deltapix_y = 1;
    if (r.npix_x > r.npix_y) deltapix_y = 1 - (r.npix_x / 2) + (r.npix_y / 2);
    else if (r.npix_x < r.npix_y) deltapix_x = 1 - (r.npix_y / 2) + (r.npix_x / 2);
    const double size_pix = r.map_size > 0.0 ? maxigrid / r.map_size : 0.0;
    double x1 = p.x, y1 = p.y, z1 = p.z;
    double u1 = p.u, v1 = p.v, w1 = p.w;
    double stok[4] = {p.S[0], p.S[1], p.S[2], p.S[3]};
    if (w1 < 0.0) {
      if (r.l_sym_centrale) { x1 = -x1; y1 = -y1; z1 = -z1; u1 = -u1; v1 = -v1; w1 = -w1; stok[2] = -stok[2]; }
      else return 0;
    }
    int capt = (int)((-1.0 * w1 + 1.0) * r.N_thet) + 1;
    if (capt == (r.N_thet + 1)) capt = r.N_thet;
    if (r.lorigine && capt == r.capt_interet) {                        // :348-357
      if (p.flag_star) t.star_origin[lambda - 1] += stok[0];
      else t.disk_origin[(size_t)(lambda - 1) + (size_t)o.n_lambda * (p.icell - 1)] += stok[0];
    }
    if (r.lmono0 && !r.loutput_mc) return capt;                        // :360
    if (r.lonly_capt_interet) { if ((capt > r.capt_sup) || (capt < r.capt_inf)) return capt; }
    int c_phi;
    if (r.l_sym_axiale) {
      if (v1 < 0.0) { v1 = -v1; y1 = -y1; stok[2] = -stok[2]; }
      if (w1 == 1.0) c_phi = 1; else c_phi = (int)(std::atan2(v1, u1) / pi * r.N_phi) + 1;
    } else {
      if (w1 == 1.0) c_phi = 1; else c_phi = (int)(fmodulo(std::atan2(u1, v1) + pi / 2, 2 * pi) / (2 * pi) * r.N_phi) + 1;
    }
    if (c_phi == (r.N_phi + 1)) c_phi = r.N_phi; else if (c_phi == 0) c_phi = 1;
    if (r.lmono0) {                                                     // map creation :396-570
      double xprim, yprim, zprim;
      rotation(x1, y1, z1, u1, v1, w1, xprim, yprim, zprim);
      double ytmp = yprim, ztmp = zprim;
      yprim = ytmp * r.cos_disk + ztmp * r.sin_disk;
      zprim = ztmp * r.cos_disk - ytmp * r.sin_disk;
      const double zoom = (double)r.zoom;
      const int imap1 = (int)((yprim * zoom + 0.5 * r.map_size) * size_pix) + deltapix_x;
      if (imap1 <= 0 || imap1 > r.npix_x) return capt;
      const int jmap1 = (int)((zprim * zoom + 0.5 * r.map_size) * size_pix) + deltapix_y;
      if (jmap1 <= 0 || jmap1 > r.npix_y) return capt;
      const size_t plane = (size_t)r.npix_x * r.npix_y * r.N_thet * r.N_phi;
      auto pix = [&](int im, int jm) { return (size_t)(im - 1) + (size_t)r.npix_x * ((size_t)(jm - 1) + (size_t)r.npix_y * ((size_t)(capt - 1) + (size_t)r.N_thet * (c_phi - 1))); };
      const int i_contrib = n_Stokes + (p.flag_star ? (p.flag_scatt ? 1 : 0) : (p.flag_scatt ? 3 : 2));
      auto add = [&](size_t q, double f, double sign_u) {
        t.stokes_map[q] += f * stok[0];
        if (r.lsepar_pola) { t.stokes_map[plane + q] += f * stok[1]; t.stokes_map[2 * plane + q] += sign_u * f * stok[2]; t.stokes_map[3 * plane + q] += f * stok[3]; }
        if (r.lsepar_contrib) t.stokes_map[(size_t)i_contrib * plane + q] += f * stok[0];
      };
      if (r.l_sym_ima) {
        add(pix(imap1, jmap1), 0.5, 1.0);
        ytmp = -ytmp;                                                   // mirror photon :468-471
        yprim = ytmp * r.cos_disk - ztmp * r.sin_disk;
        zprim = ztmp * r.cos_disk + ytmp * r.sin_disk;
        const int imap2 = (int)((yprim * zoom + 0.5 * r.map_size) * size_pix) + deltapix_x;
        if (imap2 <= 0 || imap2 > r.npix_x) return capt;
        const int jmap2 = (int)((zprim * zoom + 0.5 * r.map_size) * size_pix) + deltapix_y;
        if (jmap2 <= 0 || jmap2 > r.npix_y) return capt;
        if ((imap1 == imap2) && (jmap1 == jmap2)) add(pix(imap1, jmap1), 0.5, 1.0);
        else add(pix(imap2, jmap2), 0.5, -1.0);                         // U changes sign in the mirror pixel :520
      } else add(pix(imap1, jmap1), 1.0, 1.0);
      return capt;
    }
    size_t ix = (size_t)(lambda - 1) + (size_t)o.n_lambda * ((size_t)(capt - 1) + (size_t)r.N_thet * (c_phi - 1));
    t.sed[0][ix] += stok[0]; t.sed[1][ix] += stok[1]; t.sed[2][ix] += stok[2]; t.sed[3][ix] += stok[3];
    t.sed[4][ix] += 1.0;
    if (p.flag_star) { if (p.flag_scatt) t.sed[6][ix] += stok[0]; else t.sed[5][ix] += stok[0]; }
    else { if (p.flag_scatt) t.sed[8][ix] += stok[0]; else t.sed[7][ix] += stok[0]; }
    return capt;
  }

  // =====================================================================
  // dust_transfer.f90:439-572  mc_photon_loop
  // =====================================================================
  void alloc_tallies(int nthreads, bool reset) {
    nb_proc = nthreads;
    n_Stokes = r.lsepar_pola ? 4 : 1;
    N_type_flux = n_Stokes + (r.lsepar_contrib ? 4 : 0);       // init_mcfost.f90:1604-1616
    bool need_xJ = (r.letape_th && r.lxJ_abs_step1) || (!r.letape_th && r.lxJ_abs);
    bool need_xI = (!r.letape_th) && r.lscatt_ray_tracing1;
    size_t nsed = (size_t)o.n_lambda * r.N_thet * r.N_phi;
    size_t nxI = need_xI ? (size_t)n_az_rt * 2 * N_type_flux * r.RT_n_incl * r.RT_n_az * g.n_cells : 0;
    size_t nrt = (size_t)std::max(1, r.RT_n_incl * r.RT_n_az);
    const bool need_I = (!r.letape_th) && r.lscatt_ray_tracing2;
    size_t nI = need_I ? (size_t)N_type_flux * r.n_theta_I * r.n_phi_I * g.n_cells : 0;
    if ((int)T.size() != nthreads) { T.clear(); T.resize(nthreads); reset = true; }
    for (auto& t : T) {
      if (reset || t.xKJ_abs.size() != (size_t)g.n_cells) {
        t.xKJ_abs.assign(g.n_cells, 0.0);
        t.xT_ech.assign(g.n_cells, 2);                 // thermal_emission.f90:119, 2164
        t.xT_ech_1grain.assign(has_gr && r.lRE_nLTE ? (size_t)nk_nLTE() * g.n_cells : 0, 2);      // :165
        t.xT_ech_1grain_nRE.assign(has_gr && r.lnRE ? (size_t)nk_nRE() * g.n_cells : 0, 2);       // :188
        t.E_abs_nRE = 0.0;
        t.stokes_map.assign((r.lmono0 && r.loutput_mc) ? (size_t)r.npix_x * r.npix_y * r.N_thet * r.N_phi * N_type_flux : 0, 0.0);
        t.xN_abs.assign(r.lxN_abs ? (size_t)g.n_cells * (r.letape_th ? 1 : o.n_lambda) : 0, 0.0);
        t.star_origin.assign(r.lorigine ? o.n_lambda : 0, 0.0);
        t.disk_origin.assign(r.lorigine ? (size_t)o.n_lambda * g.n_cells : 0, 0.0);
        t.n_phot_envoyes.assign(o.n_lambda, 0.0);
        for (auto& s : t.sed) s.assign(nsed, 0.0);
        t.xJ_abs.assign(need_xJ ? (size_t)g.n_cells * o.n_lambda : 0, 0.0);
        t.xI_scatt.assign(nxI, 0.0f);
        t.I_spec.assign(nI, 0.0f); t.I_spec_star.assign(need_I ? g.n_cells : 0, 0.0f);
        for (double& s : t.stats) s = 0;
      }
      if (t.xJ_abs.size() != (need_xJ ? (size_t)g.n_cells * o.n_lambda : 0)) t.xJ_abs.assign(need_xJ ? (size_t)g.n_cells * o.n_lambda : 0, 0.0);
      if (t.xI_scatt.size() != nxI) t.xI_scatt.assign(nxI, 0.0f);
      if (t.I_spec.size() != nI) { t.I_spec.assign(nI, 0.0f); t.I_spec_star.assign(need_I ? g.n_cells : 0, 0.0f); }
      if (t.sed[0].size() != nsed) for (auto& s : t.sed) s.assign(nsed, 0.0);
      t.itheta_rt1.assign(nrt, 1); t.cos_omega_rt1.assign(nrt, 0.0); t.sin_omega_rt1.assign(nrt, 0.0);
    }
  }

  int mc_photon_loop(int nthreads, const double* rec, int64_t n_rec) {
    const int lambda_in = r.lambda_in, p_lambda_in = r.p_lambda_in;
    const int n_photons2_local = r.n_photons2;
    const float n_phot_lim = r.n_phot_lim;
    const int n_ranks = std::max(1, r.n_ranks);
#ifdef _OPENMP
    if (nthreads < 1) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    alloc_tallies(nthreads, r.reset_tallies != 0);
    for (auto& t : T) t.E_abs_nRE = 0.0;                              // :505
#pragma omp parallel num_threads(nthreads)
    {
      int id = 0;
#ifdef _OPENMP
      id = omp_get_thread_num();
#endif
      ThreadTallies& t = T[id];
      int lambda_local = lambda_in;
      const int p_lambda_local = p_lambda_in;
      PacketRng rng; std::memset(&rng, 0, sizeof rng);
      rng.rec = rec; rng.n_rec = n_rec; rng.i_rec = 0;
#pragma omp for schedule(dynamic, 1)
      for (int nnfot1 = r.nnfot1_start; nnfot1 <= r.n_photons_loop; ++nnfot1) {
        if (((nnfot1 - 1) % n_ranks) != r.rank) continue;       // multi-GPU style chunk partition
        double nnfot2 = 0.0, n_phot_envoyes_in_loop = 0.0, n_phot_sed2 = 0.0;
        const bool count_sent = (r.letape_th || r.lmono0 || r.lcount_sent) && !r.lISM_loop;        // :503-518; lISM_loop: the side loop of :941-985 counts packets that enter the model
        for (;;) {
          double p_nnfot2 = count_sent ? nnfot2 : n_phot_sed2;
          if (!((p_nnfot2 < (double)n_photons2_local) && (n_phot_envoyes_in_loop < (double)n_phot_lim))) break;
          uint64_t packet = ((uint64_t)(nnfot1 - 1) << 40) + (uint64_t)nnfot2;
          if (!rec) rng_seed_packet(&rng, r.seed, r.call_index, packet);
          nnfot2 += 1.0;
          t.n_phot_envoyes[lambda_local - 1] += 1.0;            // :531 (previous packet's lambda in thermal mode)
          n_phot_envoyes_in_loop += 1.0;
          t.stats[0] += 1;
          const double ev0 = t.stats[1] + t.stats[2];
          Packet p;
          if (!r.lmono) { float rand = (float)rng_next(&rng); select_wl_em(rand, lambda_local); }
          p.lambda = lambda_local;
          if (r.lISM_loop) {      // dust_transfer.f90:967-970
            (void)rng_next(&rng);      // (the source-choice draw of emit_packet: keeps the ISM draws on the words the device uses)
            emit_packet_ISM(rng, p.icell, p.x, p.y, p.z, p.u, p.v, p.w, p.S, p.lintersect);
            p.flag_ISM = true; p.flag_star = false;
            if (p.lintersect) n_phot_sed2 += 1.0;
          } else emit_packet(rng, p);
          p.alive = true; p.flag_scatt = false;
          if (p.lintersect) { propagate_packet(rng, t, lambda_local, p_lambda_local, p); }
          if (p.alive && (!p.flag_ISM)) {
            int capt = capteur(t, lambda_local, p);
            if (capt == r.capt_sup) n_phot_sed2 += 1.0;
            t.stats[6] += 1;
          } else if (!p.alive) t.stats[5] += 1;
          logev(0);
          if (ev_out && count_sent) ev_out[(size_t)(nnfot1 - 1) * n_photons2_local + (size_t)(nnfot2 - 1.0)] = t.stats[1] + t.stats[2] - ev0;
        }
      }
    }
    return MCB_OK;
  }

  // =====================================================================
  // thermal_emission.f90:870-906 Temp_finale (Temp_LTE with id = 0: sum / minval over the id slices)
  // =====================================================================
  void Temp_finale(float* Tdust) {
    for (int icell = 1; icell <= g.n_cells; ++icell) {
      const int p_icell = lvariable_dust() ? icell : 1;
      double sum = 0.0; for (auto& t : T) sum += t.xKJ_abs[icell - 1];
      const double Qheat = sum * e.L_packet_th / volume(icell);
      float Temp;
      if (Qheat < tiny_dp) Temp = o.T_min;
      else {
        const double log_Qheat = std::log(Qheat);
        if (log_Qheat < log_Qcool(1, p_icell)) Temp = o.T_min;
        else {
          int Ti = T[0].xT_ech[icell - 1]; for (auto& t : T) Ti = std::min(Ti, t.xT_ech[icell - 1]);
          while ((log_Qcool(Ti, p_icell) < log_Qheat) && (Ti < o.n_T)) Ti = Ti + 1;
          const double frac = (log_Qheat - log_Qcool(Ti - 1, p_icell)) / (log_Qcool(Ti, p_icell) - log_Qcool(Ti - 1, p_icell));
          Temp = (float)std::exp((double)std::log(tab_Temp(Ti)) * frac + (double)std::log(tab_Temp(Ti - 1)) * (1.0 - frac));
        }
      }
      Tdust[icell - 1] = Temp;
    }
  }
  // =====================================================================
  // thermal_emission.f90:932-1014 Temp_finale_nLTE
  // =====================================================================
  void Temp_finale_nLTE(float* T1g) {
    const int nk = nk_nLTE();
    for (int icell = 1; icell <= g.n_cells; ++icell)
      for (int k = gr.grain_RE_nLTE_start; k <= gr.grain_RE_nLTE_end; ++k) {
        const int p_k = lvariable_dust() ? k : gr.grain_zone[k - 1];
        float& out = T1g[(size_t)(k - gr.grain_RE_nLTE_start) + (size_t)nk * (icell - 1)];
        if (!(dust_density_o_n_grains(p_k, icell) > tiny_dp)) { out = 0.0f; continue; }
        double J = 0.0;
        for (int l = 1; l <= o.n_lambda; ++l) J = J + gr.C_abs_norm[gl_idx(k, l)] * (xJ_abs_sum(icell, l) + gr.J0[cl_idx(icell, l)]);
        J = J * e.L_packet_th / volume(icell);
        if (J < tiny_dp) { out = o.T_min; continue; }
        const double log_E_abs = std::log(J);
        if (log_E_abs < log_E_em_1grain(k, 1)) { out = o.T_min; continue; }
        const size_t ix = (size_t)(k - gr.grain_RE_nLTE_start) + (size_t)nk * (icell - 1);
        int T_int = T[0].xT_ech_1grain[ix]; for (auto& t : T) T_int = std::max(T_int, t.xT_ech_1grain[ix]);
        while ((log_E_em_1grain(k, T_int) < log_E_abs) && (T_int < o.n_T)) T_int = T_int + 1;
        const int T2 = T_int, T1 = T_int - 1;
        const double Temp2 = tab_Temp(T2), Temp1 = tab_Temp(T1);
        const double frac = (log_E_abs - log_E_em_1grain(k, T1)) / (log_E_em_1grain(k, T2) - log_E_em_1grain(k, T1));
        out = (float)std::exp(std::log(Temp2) * frac + std::log(Temp1) * (1.0 - frac));
      }
  }

  // merge the id slices (the reference's readers do sum(..., dim=id))
  void collect(mcb_tallies* out) {
    if (!out) return;
    const size_t nc = g.n_cells, nl = o.n_lambda;
    if (out->xKJ_abs) { for (size_t i = 0; i < nc; ++i) { double s = 0; for (auto& t : T) s += t.xKJ_abs[i]; out->xKJ_abs[i] = s; } }
    if (out->xJ_abs && !T.empty() && !T[0].xJ_abs.empty()) { for (size_t i = 0; i < nc * nl; ++i) { double s = 0; for (auto& t : T) s += t.xJ_abs[i]; out->xJ_abs[i] = s; } }
    if (out->xT_ech) { for (size_t i = 0; i < nc; ++i) { int m = T[0].xT_ech[i]; for (auto& t : T) m = std::max(m, t.xT_ech[i]); out->xT_ech[i] = m; } }
    if (out->xT_ech_1grain) for (size_t i = 0; i < T[0].xT_ech_1grain.size(); ++i) { int m = 0; for (auto& t : T) m = std::max(m, t.xT_ech_1grain[i]); out->xT_ech_1grain[i] = m; }
    if (out->xT_ech_1grain_nRE) for (size_t i = 0; i < T[0].xT_ech_1grain_nRE.size(); ++i) { int m = 0; for (auto& t : T) m = std::max(m, t.xT_ech_1grain_nRE[i]); out->xT_ech_1grain_nRE[i] = m; }
    if (out->E_abs_nRE) { double s = 0; for (auto& t : T) s += t.E_abs_nRE; *out->E_abs_nRE = s; }
    if (out->stokes_map) for (size_t i = 0; i < T[0].stokes_map.size(); ++i) { double s = 0; for (auto& t : T) s += t.stokes_map[i]; out->stokes_map[i] = s; }
    if (out->xN_abs) for (size_t i = 0; i < T[0].xN_abs.size(); ++i) { double s = 0; for (auto& t : T) s += t.xN_abs[i]; out->xN_abs[i] = s; }
    if (out->star_origin) for (size_t i = 0; i < T[0].star_origin.size(); ++i) { double s = 0; for (auto& t : T) s += t.star_origin[i]; out->star_origin[i] = s; }
    if (out->disk_origin) for (size_t i = 0; i < T[0].disk_origin.size(); ++i) { double s = 0; for (auto& t : T) s += t.disk_origin[i]; out->disk_origin[i] = s; }
    if (out->n_phot_envoyes) { for (size_t i = 0; i < nl; ++i) { double s = 0; for (auto& t : T) s += t.n_phot_envoyes[i]; out->n_phot_envoyes[i] = s; } }
    double* sp[9] = {out->sed, out->sed_q, out->sed_u, out->sed_v, out->n_phot_sed, out->sed_star, out->sed_star_scat, out->sed_disk, out->sed_disk_scat};
    for (int a = 0; a < 9; ++a) if (sp[a]) { size_t n = T[0].sed[a].size(); for (size_t i = 0; i < n; ++i) { double s = 0; for (auto& t : T) s += t.sed[a][i]; sp[a][i] = s; } }
    if (out->xI_scatt && !T[0].xI_scatt.empty()) {
      size_t n = T[0].xI_scatt.size();
      for (size_t i = 0; i < n; ++i) { float s = 0; for (auto& t : T) s += t.xI_scatt[i]; out->xI_scatt[i] = s; }   // sum(xI_scatt(...,:)) fp32, dust_ray_tracing.f90:689
      out->N_type_flux = N_type_flux;
    }
    if (out->I_spec && !T[0].I_spec.empty()) {
      for (size_t i = 0; i < T[0].I_spec.size(); ++i) { float s = 0; for (auto& t : T) s += t.I_spec[i]; out->I_spec[i] = s; }
      for (size_t i = 0; i < T[0].I_spec_star.size(); ++i) { float s = 0; for (auto& t : T) s += t.I_spec_star[i]; out->I_spec_star[i] = s; }
      out->N_type_flux = N_type_flux;
    }
    if (out->stats) for (int a = 0; a < 12; ++a) { double s = 0; for (auto& t : T) s += t.stats[a]; out->stats[a] = s; }
  }
};

}  // namespace

// ==========================================================================
// C API (ctypes).  Names are oracle_* so they can never be mistaken for the
// product's mcfost_b200_* entry points.
// ==========================================================================
extern "C" {

void* oracle_create() { return new Oracle(); }
void oracle_destroy(void* h) { delete (Oracle*)h; }
const char* oracle_last_error(void* h) { return ((Oracle*)h)->err; }

int oracle_set_grid(void* h, const mcb_grid* g) {
  Oracle* O = (Oracle*)h;
  O->g = *g;
  O->has_grid = true;
  O->ntot2 = g->n_cells;
  if (g->kind == MCB_GRID_CYL || g->kind == MCB_GRID_SPH) {
    int rc = O->build_cell_mapping();
    if (rc) return rc;
    if (g->cell_map_i && g->n_cells_tot == O->ntot2)
      for (int ic = 1; ic <= O->ntot2; ++ic)
        if (g->cell_map_i[ic - 1] != O->cmap_i[ic] || g->cell_map_j[ic - 1] != O->cmap_j[ic] || g->cell_map_k[ic - 1] != O->cmap_k[ic]) return MCB_ERR_CELL_MAP;
  }
  O->dark.assign((size_t)O->ntot2 + 1, 0);
  return MCB_OK;
}
int oracle_n_cells_tot(void* h) { return ((Oracle*)h)->ntot2; }
int oracle_get_cell_maps(void* h, int32_t* ci, int32_t* cj, int32_t* ck, int32_t* lexit) {
  Oracle* O = (Oracle*)h;
  for (int ic = 1; ic <= O->ntot2; ++ic) { ci[ic - 1] = O->cmap_i[ic]; cj[ic - 1] = O->cmap_j[ic]; ck[ic - 1] = O->cmap_k[ic]; lexit[ic - 1] = O->lexit[ic]; }
  return MCB_OK;
}
int oracle_set_dark_zone(void* h, const int32_t* dz) {
  Oracle* O = (Oracle*)h;
  O->dark.assign((size_t)O->ntot2 + 1, 0);
  if (dz) for (int i = 1; i <= O->g.n_cells; ++i) O->dark[i] = dz[i - 1] != 0;
  return MCB_OK;
}
int oracle_set_opacity(void* h, const mcb_opacity* o) { Oracle* O = (Oracle*)h; O->o = *o; O->has_op = true; return MCB_OK; }
int oracle_set_emission(void* h, const mcb_emission* e) { Oracle* O = (Oracle*)h; O->e = *e; O->has_em = true; return MCB_OK; }

int oracle_set_grains(void* h, const mcb_grains* g) { Oracle* O = (Oracle*)h; O->gr = *g; O->has_gr = true; return MCB_OK; }

static int check_run(Oracle* O, const mcb_run_params* r) {
  if (!O->has_grid || !O->has_op || !O->has_em) { snprintf(O->err, sizeof O->err, "run before uploads"); return MCB_ERR_STATE; }
  if ((r->lscattering_method1 || !r->lonly_LTE || r->low_mem_th_emission) && !O->has_gr) { snprintf(O->err, sizeof O->err, "per-grain mode without grain tables"); return MCB_ERR_STATE; }
  if (r->lmono0 && r->loutput_mc && (r->npix_x < 1 || r->npix_y < 1 || !(r->map_size > 0.0))) { snprintf(O->err, sizeof O->err, "loutput_mc needs npix_x, npix_y, map_size"); return MCB_ERR_BAD_ARG; }
  if ( (r->lscatt_ray_tracing2 && O->g.l3D)) { snprintf(O->err, sizeof O->err, "mode not built in the oracle"); return MCB_ERR_UNSUPPORTED; }
  return MCB_OK;
}

// n_threads <= 0: all OpenMP threads. rec/n_rec: optional recorded RNG stream (single thread only).
int oracle_run(void* h, const mcb_run_params* r, mcb_tallies* out, int n_threads, const double* rec, int64_t n_rec) {
  Oracle* O = (Oracle*)h;
  O->r = *r;
  int rc = check_run(O, r); if (rc) return rc;
  if (rec) n_threads = 1;
  if (r->lISM_loop) {      // dust_transfer.f90:941-945: the ray-tracing accumulators are switched off around the side loop
    if (r->letape_th) { snprintf(O->err, sizeof O->err, "lISM_loop is not a thermal-step mode"); return MCB_ERR_BAD_ARG; }
    O->r.lscatt_ray_tracing1 = 0; O->r.lscatt_ray_tracing2 = 0;
  }
  if (r->lMRW) {
    if (!r->letape_th || !r->lonly_LTE || r->low_mem_th_emission || r->lxJ_abs_step1) { snprintf(O->err, sizeof O->err, "lMRW: thermal step with lonly_LTE only"); return MCB_ERR_UNSUPPORTED; }
    if (O->zeta_tab.empty()) O->initialize_cumulative_zeta();
    O->compute_MRW_means();
  }
  rc = O->mc_photon_loop(n_threads, rec, n_rec); if (rc) return rc;
  O->collect(out);
  return MCB_OK;
}

// Temp_finale on tallies supplied by the caller (e.g. downloaded from the GPU): they become thread 0's tallies
int oracle_temp_finale_of(void* h, const double* xKJ_abs, const int32_t* xT_ech, float* Tdust) {
  Oracle* O = (Oracle*)h;
  if (!O->has_grid || !O->has_op || !O->has_em) return MCB_ERR_STATE;
  O->alloc_tallies(1, true);
  for (int i = 0; i < O->g.n_cells; ++i) { O->T[0].xKJ_abs[i] = xKJ_abs[i]; O->T[0].xT_ech[i] = xT_ech[i]; }
  O->Temp_finale(Tdust);
  return MCB_OK;
}
int oracle_temp_finale(void* h, float* Tdust) { Oracle* O = (Oracle*)h; if (O->T.empty()) return MCB_ERR_STATE; O->Temp_finale(Tdust); return MCB_OK; }
int oracle_temp_finale_nlte(void* h, float* T1g) { Oracle* O = (Oracle*)h; if (O->T.empty() || O->T[0].xT_ech_1grain.empty() || O->T[0].xJ_abs.empty()) return MCB_ERR_STATE; O->Temp_finale_nLTE(T1g); return MCB_OK; }

int oracle_cross_cell(void* h, int64_t n, const double* x0, const double* y0, const double* z0, const double* u, const double* v, const double* w,
                      const int32_t* icell, const int32_t* previous_cell, double* x1, double* y1, double* z1, int32_t* next_cell,
                      double* l, double* l_contrib, double* l_void_before) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i) {
    int nc;
    O->cross_cell(x0[i], y0[i], z0[i], u[i], v[i], w[i], icell[i], previous_cell ? previous_cell[i] : 0, x1[i], y1[i], z1[i], nc, l[i], l_contrib[i], l_void_before[i]);
    next_cell[i] = nc;
  }
  return MCB_OK;
}
int oracle_distance_to_closest_wall(void* h, int64_t n, const int32_t* icell, const double* x, const double* y, const double* z, double* s) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i) s[i] = O->distance_to_closest_wall(icell[i], x[i], y[i], z[i]);
  return MCB_OK;
}
int oracle_mrw_tables(void* h, double* A, double* B, double* Cc) {
  Oracle* O = (Oracle*)h;
  O->mrw_for = nullptr; O->compute_MRW_means();
  const size_t n = O->mrw_A.size();
  for (size_t i = 0; i < n; ++i) { A[i] = O->mrw_A[i]; B[i] = O->mrw_B[i]; Cc[i] = O->mrw_C[i]; }
  return MCB_OK;
}
int oracle_zeta_table(void* h, double* zeta, int n) {
  Oracle* O = (Oracle*)h;
  if (O->zeta_tab.empty()) O->initialize_cumulative_zeta();
  for (int i = 0; i < n && i < Oracle::n_zeta; ++i) zeta[i] = O->zeta_tab[i];
  return Oracle::n_zeta;
}
double oracle_sample_zeta(void* h, double zr) {
  Oracle* O = (Oracle*)h;
  if (O->zeta_tab.empty()) O->initialize_cumulative_zeta();
  return O->sample_zeta(zr);
}
int oracle_index_cell(void* h, int64_t n, const double* x, const double* y, const double* z, int32_t* icell) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i) { int ic = 0; O->index_cell(x[i], y[i], z[i], ic); icell[i] = ic; }
  return MCB_OK;
}
int oracle_move_to_grid(void* h, int64_t n, double* x, double* y, double* z, const double* u, const double* v, const double* w, int32_t* icell, int32_t* lintersect) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i) { int ic = 0; bool li = false; O->move_to_grid(x[i], y[i], z[i], u[i], v[i], w[i], ic, li); icell[i] = li ? ic : 0; lintersect[i] = li; }
  return MCB_OK;
}
int oracle_optical_length_tot(void* h, int64_t n, int32_t lambda, const double* x, const double* y, const double* z, const double* u, const double* v, const double* w,
                              const int32_t* icell, double* tau_tot, double* lmin, double* lmax, int32_t* n_steps) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i) { float tt; int ns; O->optical_length_tot(lambda, icell[i], x[i], y[i], z[i], u[i], v[i], w[i], tt, lmin[i], lmax[i], ns); tau_tot[i] = tt; if (n_steps) n_steps[i] = ns; }
  return MCB_OK;
}
int oracle_define_dark_zone(void* h, int32_t lambda, float tau_max, const double* r_grid, const double* z_grid, int32_t n_regions,
                            const int32_t* iRmin, const int32_t* iRmax, const double* dust_sum, int32_t* dark, int32_t* ri_in, int32_t* ri_out,
                            int32_t* zj_sup, int32_t* zj_inf, int32_t* l_is_dark) {
  ((Oracle*)h)->define_dark_zone(lambda, tau_max, r_grid, z_grid, n_regions, iRmin, iRmax, dust_sum, dark, ri_in, ri_out, zj_sup, zj_inf, l_is_dark);
  return MCB_OK;
}
int oracle_init_reemission(void* h, const double* tab_lambda, const double* tab_delta_lambda, double* logQ, double* cdf) {
  ((Oracle*)h)->init_reemission(tab_lambda, tab_delta_lambda, logQ, cdf);
  return MCB_OK;
}
int oracle_init_reemission_grains(void* h, const double* tab_lambda, const double* tab_delta_lambda, const float* C_abs_norm, int32_t n_grains_tot,
                                  int32_t k_start, int32_t k_end, double* logE, double* E_em, double* cdf) {
  ((Oracle*)h)->init_reemission_grains(tab_lambda, tab_delta_lambda, C_abs_norm, n_grains_tot, k_start, k_end, logE, E_em, cdf);
  return MCB_OK;
}
int oracle_init_dust_source_fct1(void* h, int32_t lambda, int32_t iRT, int32_t n_RT, double photon_energy, const double* J_th, const float* xI,
                                 int32_t az_dim, int32_t th_dim, int32_t n_az_rt, int32_t n_theta_rt, int32_t N_type_flux, int32_t n_Stokes,
                                 int32_t lsepar_pola, int32_t lsepar_contrib, double* eps) {
  ((Oracle*)h)->init_dust_source_fct1(lambda, iRT, n_RT, photon_energy, J_th, xI, az_dim, th_dim, n_az_rt, n_theta_rt, N_type_flux, n_Stokes,
                                      lsepar_pola != 0, lsepar_contrib != 0, eps);
  return MCB_OK;
}
int oracle_integ_ray_dust(void* h, int32_t lambda, int64_t n, const double* x, const double* y, const double* z, const double* u, const double* v,
                          const double* w, const int32_t* icell, float tau_dark_zone_obs, const double* eps, int32_t az_dim, int32_t th_dim,
                          int32_t n_az_rt, int32_t N_type_flux, double* out) {
  Oracle* O = (Oracle*)h;
  for (int64_t i = 0; i < n; ++i)
    O->integ_ray_dust(lambda, icell[i], x[i], y[i], z[i], u[i], v[i], w[i], tau_dark_zone_obs, eps, az_dim, th_dim, n_az_rt, N_type_flux, out + (size_t)N_type_flux * i);
  return MCB_OK;
}
int oracle_repartition_energie(void* h, int32_t lambda_first, int32_t lambda_last, const float* Tdust, const double* tab_lambda, const double* E_stars,
                               const double* E_ISM, const double* weight, double* E_disk, double* frac_E_stars, double* frac_E_disk,
                               double* weight_norm, double* prob_E_cell) {
  Oracle* O = (Oracle*)h;
  const size_t n1 = (size_t)O->g.n_cells + 1;
  for (int l = lambda_first; l <= lambda_last; ++l) {
    double wn = 0.0;
    if (!O->repartition_energie(l, Tdust, tab_lambda, E_stars[l - 1], E_ISM ? E_ISM[l - 1] : 0.0, weight, prob_E_cell + n1 * (size_t)(l - 1),
                                E_disk[l - 1], frac_E_stars[l - 1], frac_E_disk[l - 1], &wn)) return MCB_ERR_BAD_ARG;
    if (weight_norm) weight_norm[l - 1] = wn;
  }
  return MCB_OK;
}
int oracle_compute_column(void* h, int32_t lambda, const double* factor, const double* cx, const double* cy, const double* cz, float* column) {
  ((Oracle*)h)->compute_column(lambda, factor, cx, cy, cz, column);
  return MCB_OK;
}
int oracle_physical_length(void* h, int64_t n, int32_t lambda, double* x, double* y, double* z, double* u, double* v, double* w,
                           int32_t* icell, const float* tau, float* ltot, int32_t* flag_sortie, int32_t* lpacket_alive) {
  Oracle* O = (Oracle*)h;
  ThreadTallies t; for (double& s : t.stats) s = 0;
  double S0[4] = {0, 0, 0, 0};
  for (int64_t i = 0; i < n; ++i) {
    int ic = icell[i]; bool fs = false, alive = true; float lt = 0;
    O->physical_length(t, lambda, 1, S0, ic, x[i], y[i], z[i], u[i], v[i], w[i], false, false, tau[i], lt, fs, alive, false);
    icell[i] = ic; ltot[i] = lt; flag_sortie[i] = fs; lpacket_alive[i] = alive;
  }
  return MCB_OK;
}

// Stand-alone helpers exposed for unit tests of the samplers.
void oracle_cdapres(double cospsi, double phi, double u0, double v0, double w0, double* out3) { Oracle::cdapres(cospsi, phi, u0, v0, w0, out3[0], out3[1], out3[2]); }
void oracle_rotation(double xi, double yi, double zi, double u1, double v1, double w1, double* out3) { Oracle::rotation(xi, yi, zi, u1, v1, w1, out3[0], out3[1], out3[2]); }
void oracle_hg(float g, float rand, int32_t* itheta, double* cospsi) { int it; Oracle::hg(g, rand, it, *cospsi); *itheta = it; }
void oracle_update_stokes(double* S, double u0, double v0, double w0, double u1, double v1, double w1, const double* M16) {
  double M[4][4]; for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) M[a][b] = M16[a * 4 + b];
  Oracle::update_Stokes(S, u0, v0, w0, u1, v1, w1, M);
}
void oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void oracle_rng_stream(uint64_t seed, uint32_t call_index, uint64_t packet, int n, double* out) {
  PacketRng g; std::memset(&g, 0, sizeof g); rng_seed_packet(&g, seed, call_index, packet);
  for (int i = 0; i < n; ++i) out[i] = rng_next(&g);
}
// instrumentation: per-packet (cell steps + interactions), thermal mode, length n_photons_loop*n_photons2
void oracle_set_event_buffer(void* h, double* buf) { ((Oracle*)h)->ev_out = buf; }
int64_t oracle_set_event_log(void* h, uint8_t* buf, int64_t cap) { Oracle* O = (Oracle*)h; int64_t n = O->ev_log_n; O->ev_log = buf; O->ev_log_cap = cap; O->ev_log_n = 0; return n; }
int64_t oracle_event_log_size(void* h) { return ((Oracle*)h)->ev_log_n; }
int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
