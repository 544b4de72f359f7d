// Packet physics on the device: emission, optical-depth walk, scattering,
// Bjorkman & Wood re-emission, detectors, and the persistent photon-loop kernel.
//
// Reference routines covered (paths relative to reference src/):
//   mc_photon_loop dust_transfer.f90:439-572, emit_packet :1047-1151,
//   propagate_packet :1155-1409, physical_length optical_depth.f90:21-182,
//   save_radiation_field radiation_field.f90:31-135, angles_scatt_rt1 /
//   calc_xI_scatt[_pola] dust_ray_tracing.f90:409-632, Temp_LTE /
//   im_reemission_LTE thermal_emission.f90:649-771, select_wl_em :364,
//   select_cellule :2044, angle_diff_theta_pos / hg / update_Stokes /
//   get_Mueller_matrix_per_cell scattering.f90:1187-1475, cdapres / rotation
//   utils.f90:553-599,1636-1688, select_star / emit_packet_uniform_sphere /
//   emit_packet_ISM / intersect_stars stars.f90:75-169,728-884, capteur
//   output.f90:294-595 (SED branch).
//
// Execution model (B200): one packet per thread, persistent warps.  Each warp
// iteration runs the phases FETCH -> TAU -> FLY -> INTERACT under warp-uniform
// guards so that lanes in the same phase execute together; new packets are
// claimed from a global counter with one warp-aggregated atomic.  The small hot
// tables (radial / vertical walls, kappa(lambda), albedo, log Qcool, the k dB/dT
// CDF, the s11 CDF, cos table, emission spectra) are staged once per block in
// shared memory (SmemLayout, ~48 KB for ref4.1) when the dust is not cell-
// dependent; per-cell arrays stay in global memory behind L1/L2.  Tallies are L2
// atomics (red.global.add.f64); the running cell temperature reads them back
// with ld.global.cg (L1 is not coherent with L2 atomics).  The next-cell half of
// a crossing is skipped when the flight ends inside the cell.
#pragma once
#include "model.cuh"
#include "philox.cuh"
#include "geom_rz.cuh"
#include "geom_vor.cuh"

namespace mcb {

enum { ST_FETCH = 0, ST_TAU = 1, ST_FLY = 2, ST_INTERACT = 3, ST_DONE = 4 };
enum { STAT_PACKETS = 0, STAT_STEPS, STAT_INTERACT, STAT_SCATT, STAT_ABS, STAT_KILLED, STAT_ESCAPED, STAT_BOUNCE };

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

// ---- opacity / thermal table accessors (SM: shared-memory staging, p_n_cells == 1) ----
__device__ __forceinline__ const float* smf(int word_off) { return reinterpret_cast<const float*>(smd() + word_off); }
template <bool SM> __device__ __forceinline__ double t_kappa(const DevModel& m, int p_icell, int lambda) {
  return SM ? smd()[m.sm.kappa + lambda - 1] : __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ double t_kappa_abs(const DevModel& m, int p_icell, int lambda) {
  return SM ? smd()[m.sm.kappa_abs + lambda - 1] : __ldg(m.kappa_abs + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ float t_albedo(const DevModel& m, int p_icell, int lambda) {
  return SM ? smf(m.sm.albedo)[lambda - 1] : __ldg(m.albedo + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ float t_gfac(const DevModel& m, int p_icell, int lambda) {
  return SM ? smf(m.sm.gfac)[lambda - 1] : __ldg(m.gfac + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ double t_logQ(const DevModel& m, int t, int p_icell) {      // t 1-based
  return SM ? smd()[m.sm.logQ + t - 1] : __ldg(m.logQ + (size_t)m.n_T * (p_icell - 1) + t - 1); }
template <bool SM> __device__ __forceinline__ double t_kdB(const DevModel& m, int l, int t, int p_icell) {  // l, t 1-based
  return SM ? smd()[m.sm.kdB + (l - 1) + m.n_lambda * (t - 1)]
            : __ldg(m.kdB + (size_t)m.n_lambda * ((t - 1) + (size_t)m.n_T * (p_icell - 1)) + (l - 1)); }
template <bool SM> __device__ __forceinline__ double t_cos(const DevModel& m, int k) { return SM ? smd()[m.sm.cos_tab + k] : __ldg(m.cos_tab + k); }
template <bool SM> __device__ __forceinline__ float t_prob_s11(const DevModel& m, int k, int p_icell, int p_lambda) {
  return SM ? smf(m.sm.prob_s11)[k] : __ldg(m.prob_s11 + (size_t)(NANG + 1) * ((p_icell - 1) + (size_t)m.p_n_cells * (p_lambda - 1)) + k); }
template <bool SM> __device__ __forceinline__ double t_spec_cumul(const DevModel& m, int k) { return SM ? smd()[m.sm.spec_cumul + k] : __ldg(m.spec_cumul + k); }
template <bool SM> __device__ __forceinline__ double t_frac_star(const DevModel& m, int lambda) { return SM ? smd()[m.sm.frac_star + lambda - 1] : __ldg(m.frac_star + lambda - 1); }
template <bool SM> __device__ __forceinline__ double t_frac_disk(const DevModel& m, int lambda) { return SM ? smd()[m.sm.frac_disk + lambda - 1] : __ldg(m.frac_disk + lambda - 1); }

// one block-wide copy of the staged tables into shared memory
__device__ __forceinline__ void stage_tables(const DevModel& m, int p_lambda_in) {
  double* sd = reinterpret_cast<double*>(mcb_smem_raw);
  const SmemLayout& L = m.sm;
  auto cp = [&](int off, const double* src, int n) { if (src) for (int i = threadIdx.x; i < n; i += blockDim.x) sd[off + i] = src[i]; };
  auto cpf = [&](int off, const float* src, int n) { if (src) { float* d = reinterpret_cast<float*>(sd + off); for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = src[i]; } };
  cp(L.r_lim_2, m.r_lim_2, m.n_rad + 1);
  if (m.kind == 1) { cp(L.zmax, m.zmax, m.n_rad); if (m.z_regular) cp(L.zl, m.cell_height, m.n_rad); else cp(L.zl, m.z_lim, m.n_rad * (m.nz + 2)); }
  else cp(L.tan_theta, m.tan_theta_lim, m.nz + 1);
  if (m.l3D) cp(L.tan_phi, m.tan_phi_lim, m.n_az);
  cp(L.kappa, m.kappa, m.n_lambda); cp(L.kappa_abs, m.kappa_abs, m.n_lambda);
  cpf(L.albedo, m.albedo, m.n_lambda); cpf(L.gfac, m.gfac, m.n_lambda);
  cp(L.logQ, m.logQ, m.n_T); cp(L.kdB, m.kdB, m.n_lambda * m.n_T);
  cp(L.cos_tab, m.cos_tab, NANG + 1);
  if (m.prob_s11) cpf(L.prob_s11, m.prob_s11 + (size_t)(NANG + 1) * (p_lambda_in - 1), NANG + 1);
  cp(L.spec_cumul, m.spec_cumul, m.n_lambda + 1); cp(L.frac_star, m.frac_star, m.n_lambda); cp(L.frac_disk, m.frac_disk, m.n_lambda);
  __syncthreads();
}

// ---- utils.f90:1636-1688 cdapres ------------------------------------------
__device__ __forceinline__ void cdapres(double cospsi, double sphi, double cphi, double u0, double v0, double w0,
                                        double& u1, double& v1, double& w1) {
  double spsi = sqrt(1.0 - cospsi * cospsi);
  double a = spsi * cphi, b = spsi * sphi;
  if (fabs(w0) <= (double)0.999999f) {
    double c = sqrt(1.0 - w0 * w0), cm1 = 1.0 / c, aw0 = a * w0;
    u1 = (aw0 * u0 - b * v0) * cm1 + cospsi * u0;
    v1 = (aw0 * v0 + b * u0) * cm1 + cospsi * v0;
    w1 = cospsi * w0 - a * c;
  } else { u1 = a; v1 = b; w1 = cospsi; }
}
// ---- utils.f90:553-599 rotation -------------------------------------------
__device__ __forceinline__ void rotation(double xi, double yi, double zi, double u1, double v1, double w1,
                                         double& xf, double& yf, double& zf) {
  double cost, sint, sing;
  if (w1 > 0.999999999) { cost = 1.0; sint = 0.0; sing = 0.0; }
  else if (fabs(u1) < MCB_TINY_REAL) { cost = 0.0; sint = 1.0; sing = sqrt(1.0 - w1 * w1); }
  else { double th = atan2(v1, u1); sincos(th, &sint, &cost); sing = sqrt(1.0 - w1 * w1); }
  double prod = cost * xi + sint * yi;
  xf = sing * prod + w1 * zi;
  yf = cost * yi - sint * xi;
  zf = sing * zi - w1 * prod;
}
// ---- random_numbers.f90:32-51 ---------------------------------------------
__device__ __forceinline__ void random_isotropic_direction(Rng& rng, double& u, double& v, double& w) {
  float rand = rng.nextf();
  w = 2.0 * rand - 1.0;
  double uv = sqrt(1.0 - w * w);
  rand = rng.nextf();
  double sp, cp;
  sincospi(2.0 * rand - 1.0, &sp, &cp);       // phi = pi*(2 rand - 1)
  u = uv * cp; v = uv * sp;
}

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

// ---- bisection samplers ------------------------------------------------------
template <bool SM>
__device__ __forceinline__ int select_wl_em(const DevModel& m, float rand) {          // thermal_emission.f90:364-400
  int kmin = 0, kmax = m.n_lambda, k = (kmin + kmax) / 2;
  while (t_spec_cumul<SM>(m, k) != (double)rand) {
    if (t_spec_cumul<SM>(m, k) < (double)rand) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
    if ((kmax - kmin) <= 1) break;
  }
  return kmax;
}
__device__ __forceinline__ int select_star(const DevModel& m, int lambda, float rand) {   // stars.f90:75-104
  int kmin = 0, kmax = m.n_stars, k = (kmax - kmin) / 2;
  while ((kmax - kmin) > 1) {
    if (__ldg(m.CDF_E_star + (lambda - 1) + m.n_lambda * k) < rand) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
  }
  return kmax;
}
__device__ __forceinline__ int select_cellule(const DevModel& m, int lambda, float rand) {   // thermal_emission.f90:2044-2074
  const double* p = m.prob_E_cell + (size_t)(m.n_cells + 1) * (lambda - 1);
  int kmin = 0, kmax = m.n_cells, k = (kmin + kmax) / 2;
  while ((kmax - kmin) > 1) {
    if (__ldg(p + k) < (double)rand) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
  }
  return kmax;
}

// ---- scattering.f90:1354-1383 hg ----------------------------------------------
__device__ __forceinline__ void hg(float g, float rand, int& itheta, double& cospsi) {
  double rand_dp = fmin((double)rand, 1.0 - 1e-6);
  if (fabsf(g) > FLT_MIN) {
    double g1 = g, g2 = g1 * g1;
    double q = (1.0 - g2) / (1.0 - g1 + 2.0 * g1 * rand_dp);
    cospsi = (1.0 + g2 - q * q) / (2.0 * g1);
  } else cospsi = 2.0 * rand_dp - 1.0;
  itheta = (int)floor(acos(cospsi) * 180.0 / MCB_PI) + 1;
  if (itheta > NANG) itheta = NANG;
}
// ---- scattering.f90:1433-1475 angle_diff_theta_pos ------------------------------
template <bool SM>
__device__ __forceinline__ void angle_diff_theta_pos(const DevModel& m, int p_lambda, int p_icell, float rand, float rand2,
                                                     int& itheta, double& cospsi) {
  int kmin = 0, kmax = NANG, k = (kmin + kmax) / 2;
  while ((kmax - kmin) > 1) {
    if (t_prob_s11<SM>(m, k, p_icell, p_lambda) < rand) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
  }
  k = kmax;
  itheta = k;
  double c0 = t_cos<SM>(m, k - 1), c1 = t_cos<SM>(m, k);
  cospsi = c0 + rand2 * (c1 - c0);
}

// ---- scattering.f90:1187-1298 update_Stokes with the per-cell Mueller matrix
// of get_Mueller_matrix_per_cell (:1328-1350); sparse products written out ------
__device__ __forceinline__ void scatter_stokes(const DevModel& m, int lambda, int itheta, float frac, int p_icell, double* S,
                                               double u0, double v0, double w0, double u1, double v1, double w1) {
  const size_t q1 = (size_t)itheta + (size_t)(NANG + 1) * ((p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)), q0 = q1 - 1;
  const float frac_m1 = 1.0f - frac;
  const double M11 = (double)1.0f;
  const double M22 = (double)__fadd_rn(__fmul_rn(__ldg(m.s22 + q1), frac), __fmul_rn(__ldg(m.s22 + q0), frac_m1));
  const double M12 = (double)__fadd_rn(__fmul_rn(__ldg(m.s12 + q1), frac), __fmul_rn(__ldg(m.s12 + q0), frac_m1));
  const double M33 = (double)__fadd_rn(__fmul_rn(__ldg(m.s33 + q1), frac), __fmul_rn(__ldg(m.s33 + q0), frac_m1));
  const double M44 = (double)__fadd_rn(__fmul_rn(__ldg(m.s44 + q1), frac), __fmul_rn(__ldg(m.s44 + q0), frac_m1));
  const double M34 = (double)__fsub_rn(__fmul_rn(-__ldg(m.s34 + q1), frac), __fmul_rn(__ldg(m.s34 + q0), frac_m1));
  const double M43 = -M34;
  double v1pi, v1pj, v1pk;
  rotation(u0, v0, w0, u1, v1, w1, v1pi, v1pj, v1pk);
  float xnyp = (float)sqrt(v1pk * v1pk + v1pj * v1pj), costhet;
  if (xnyp < 1e-10f) { xnyp = 0.0f; costhet = 1.0f; }
  else costhet = (float)(-1.0 * v1pj / (double)xnyp);
  float theta = acosf(costhet);
  if ((double)theta >= MCB_PI) theta = 0.0f;
  theta = (float)((double)theta + MCB_HALF_PI);
  float omega = 2.0f * theta;
  if (v1pk < 0.0) omega = -1.0f * omega;
  float cosw = cosf(omega), sinw = sinf(omega);
  if (fabsf(cosw) < 1e-06f) cosw = 0.0f;
  if (fabsf(sinw) < 1e-06f) sinw = 0.0f;
  const double cw = cosw, sw = sinw;
  const double S1_0 = S[0];
  // C = ROP.S ; ROP(2,2)=cw ROP(2,3)=-sw ROP(3,2)=sw ROP(3,3)=cw
  const double C0 = S[0], C1 = cw * S[1] + (-sw) * S[2], C2 = sw * S[1] + cw * S[2], C3 = S[3];
  // D = M.C
  const double D0 = M11 * C0 + M12 * C1, D1 = M12 * C0 + M22 * C1, D2 = M33 * C2 + M34 * C3, D3 = M43 * C2 + M44 * C3;
  // S = RPO.D ; RPO(2,2)=cw RPO(2,3)=sw RPO(3,2)=-sw RPO(3,3)=cw
  S[0] = D0; S[1] = cw * D1 + sw * D2; S[2] = (-sw) * D1 + cw * D2; S[3] = D3;
  if (S[0] > MCB_TINY_REAL) {
    const double S0 = S[0];
#pragma unroll
    for (int a = 0; a < 4; ++a) S[a] = S[a] * M11 * S1_0 / S0;
  }
}

// ---- thermal_emission.f90:649-771 Temp_LTE + im_reemission_LTE (high-memory
// branch): new wavelength index ------------------------------------------------
template <bool SM>
__device__ __forceinline__ int im_reemission_LTE(const DevModel& m, const DevRun& r, int idx, int p_icell, float rand2) {
  // running tally: L2-coherent load (the adds are L2 atomics)
  double Qheat = __ldcg(m.tally + m.lay.xKJ + idx) * r.nb_proc_equiv * m.L_packet_th / __ldg(m.volume + idx);
  int Ti = 2;
  double frac_T2 = 0.0;       // `frac` is left undefined by the reference at T_min; 0 chosen (same as the oracle)
  if (!(Qheat < MCB_TINY_DP)) {
    double log_Qheat = log(Qheat);
    if (!(log_Qheat < t_logQ<SM>(m, 1, p_icell))) {
      Ti = __ldcg(m.xT_ech + idx);
      while ((t_logQ<SM>(m, Ti, p_icell) < log_Qheat) && (Ti < m.n_T)) ++Ti;
      // another warp may have cached an index computed from a larger running tally: step back down
      while (Ti > 2 && !(t_logQ<SM>(m, Ti - 1, p_icell) < log_Qheat)) --Ti;
      double q1 = t_logQ<SM>(m, Ti - 1, p_icell), q2 = t_logQ<SM>(m, Ti, p_icell);
      frac_T2 = (log_Qheat - q1) / (q2 - q1);
    }
  }
  atomicMax(m.xT_ech + idx, Ti);
  const double frac_T1 = 1.0 - frac_T2;
  int l1 = 0, l2 = m.n_lambda, l = (l1 + l2) / 2;
  while ((l2 - l1) > 1) {
    double proba = frac_T1 * t_kdB<SM>(m, l, Ti - 1, p_icell) + frac_T2 * t_kdB<SM>(m, l, Ti, p_icell);
    if ((double)rand2 > proba) l1 = l; else l2 = l;
    l = (l1 + l2) / 2;
  }
  return l + 1;
}

// ---- output.f90:294-595 capteur, SED branch ------------------------------------
__device__ __forceinline__ int capteur(const DevModel& m, const DevRun& r, int lambda, double u1, double v1, double w1,
                                       const double* Sin, bool flag_star, bool flag_scatt) {
  double s0 = Sin[0], s1 = Sin[1], s2 = Sin[2], s3 = Sin[3];
  if (w1 < 0.0) {
    if (r.l_sym_centrale) { u1 = -u1; v1 = -v1; w1 = -w1; s2 = -s2; }
    else return 0;
  }
  int capt = (int)((-1.0 * w1 + 1.0) * r.N_thet) + 1;
  if (capt == r.N_thet + 1) capt = r.N_thet;
  int c_phi = 1;
  if (r.l_sym_axiale) {
    if (v1 < 0.0) { v1 = -v1; s2 = -s2; }
    if (r.N_phi > 1 && w1 != 1.0) c_phi = (int)(atan2(v1, u1) / MCB_PI * r.N_phi) + 1;
  } else {
    if (w1 != 1.0) c_phi = (int)(fmodulo(atan2(u1, v1) + MCB_PI / 2, 2 * MCB_PI) / (2 * MCB_PI) * r.N_phi) + 1;
  }
  if (c_phi == r.N_phi + 1) c_phi = r.N_phi; else if (c_phi == 0) c_phi = 1;
  const int64_t ix = (lambda - 1) + (int64_t)m.n_lambda * ((capt - 1) + (int64_t)r.N_thet * (c_phi - 1));
  double* sed = m.tally + m.lay.sed;
  const int64_t n = m.lay.n_sed;
  atomicAdd(sed + 0 * n + ix, s0);
  if (r.lsepar_pola) { atomicAdd(sed + 1 * n + ix, s1); atomicAdd(sed + 2 * n + ix, s2); atomicAdd(sed + 3 * n + ix, s3); }
  atomicAdd(sed + 4 * n + ix, 1.0);
  const int which = flag_star ? (flag_scatt ? 6 : 5) : (flag_scatt ? 8 : 7);
  atomicAdd(sed + which * n + ix, s0);
  return capt;
}

// ---- dust_ray_tracing.f90:409-476 angles_scatt_rt1 (per flight) ------------------
struct Rt1Scratch { unsigned char itheta[MAX_RT]; double cosw[MAX_RT], sinw[MAX_RT]; };

__device__ __forceinline__ void angles_scatt_rt1(const DevRun& r, double u, double v, double w, Rt1Scratch& sc) {
  for (int i = 0; i < r.n_rt; ++i) {
    float cos_scatt = (float)(r.rt_u[i] * u + r.rt_v[i] * v + r.rt_w[i] * w);
    // k = nint(acos(cos_scatt) * real(nang_scatt)/pi): fp32 product, dp division, round half away (q >= 0)
    const double q = (double)__fmul_rn(acosf(cos_scatt), (float)NANG) / MCB_PI;
    int k = (int)floor(q + 0.5);
    if (k > NANG) k = NANG;
    if (k < 1) k = 1;
    sc.itheta[i] = (unsigned char)k;
    if (r.lsepar_pola) {
      double v1pi, v1pj, v1pk;
      rotation(u, v, w, -r.rt_u[i], -r.rt_v[i], -r.rt_w[i], v1pi, v1pj, v1pk);
      double xnyp = sqrt(v1pk * v1pk + v1pj * v1pj), costhet;
      if (xnyp < (double)1e-10f) costhet = 1.0; else costhet = -1.0 * v1pj / xnyp;
      double theta = acos(costhet);
      if (theta >= MCB_PI) theta = 0.0;
      theta = theta + MCB_HALF_PI;
      double omega = 2.0 * theta;
      if (v1pk < 0.0) omega = -1.0 * omega;
      double sw, cw; sincos(omega, &sw, &cw);
      if (fabs(cw) < (double)1e-06f) cw = 0.0;
      if (fabs(sw) < (double)1e-06f) sw = 0.0;
      sc.cosw[i] = cw; sc.sinw[i] = sw;
    }
  }
}

// ---- radiation_field.f90:63-89 + dust_ray_tracing.f90:480-632: rt1 scattered
// specific intensity, fp32 atomics ---------------------------------------------------
__device__ __forceinline__ void deposit_rt1(const DevModel& m, const DevRun& r, int idx, int p_icell, int p_lambda, double l,
                                            const double* S, bool flag_star, double xm, double ym, double zm, const Rt1Scratch& sc) {
  int phi_k = 1, psup = 1;
  if (!m.l3D) {
    double phi_pos = atan2(xm, ym);
    phi_k = (int)floor(fmodulo(phi_pos, MCB_TWO_PI) / MCB_TWO_PI * N_AZ_RT) + 1;
    if (phi_k > N_AZ_RT) phi_k = N_AZ_RT;
    psup = (zm > 0.0) ? 1 : 2;
  }
  const size_t tab0 = (size_t)(NANG + 1) * ((p_icell - 1) + (size_t)m.p_n_cells * (p_lambda - 1));
  for (int i = 0; i < r.n_rt; ++i) {
    const int it = sc.itheta[i];
    // (phik, psup, itype, iRT, icell) column-major
    const size_t base = (size_t)(phi_k - 1) + (size_t)N_AZ_RT * ((size_t)(psup - 1) + 2 * ((size_t)r.n_type_flux * ((size_t)i + (size_t)r.n_rt * (size_t)idx)));
    const size_t stride = (size_t)N_AZ_RT * 2;
    const float s11 = __ldg(m.s11 + tab0 + it);
    if (!r.lsepar_pola) {
      const double flux = l * S[0] * (double)s11;
      atomicAdd(m.xI + base, (float)flux);
      if (r.lsepar_contrib) atomicAdd(m.xI + base + stride * (size_t)(r.n_stokes + (flag_star ? 2 : 4) - 1), (float)flux);
    } else {
      const float s12 = -s11 * __ldg(m.s12 + tab0 + it), s22 = s11 * __ldg(m.s22 + tab0 + it);
      const float s33 = -s11 * __ldg(m.s33 + tab0 + it), s34 = -s11 * __ldg(m.s34 + tab0 + it), s44 = -s11 * __ldg(m.s44 + tab0 + it);
      const double cw = sc.cosw[i], sw = sc.sinw[i];
      const double C0 = S[0], C1 = cw * S[1] + (-sw) * S[2], C2 = sw * S[1] + cw * S[2], C3 = S[3];
      const double D0 = (double)s11 * C0 + (double)s12 * C1, D1 = (double)s12 * C0 + (double)s22 * C1;
      const double D2 = (double)s33 * C2 + (double)(-s34) * C3, D3 = (double)s34 * C2 + (double)s44 * C3;
      const double R0 = D0, R1 = (-cw) * D1 + (-sw) * D2, R2 = (-sw) * D1 + cw * D2, R3 = D3;
      atomicAdd(m.xI + base + 0 * stride, (float)(l * R0));
      atomicAdd(m.xI + base + 1 * stride, (float)(l * R1));
      atomicAdd(m.xI + base + 2 * stride, (float)(l * R2));
      atomicAdd(m.xI + base + 3 * stride, (float)(l * R3));
      if (r.lsepar_contrib) atomicAdd(m.xI + base + stride * (size_t)(flag_star ? 5 : 7), (float)(l * R0));
    }
  }
}

// =============================================================================
// The persistent photon-loop kernel
// =============================================================================
constexpr int MC_BLOCK = 256;

template <class G, bool SM>
__global__ void __launch_bounds__(MC_BLOCK, 2)
mc_photon_loop_kernel(const __grid_constant__ DevModel m, const __grid_constant__ DevRun r) {
  using CellT = typename G::CellT;
  using Hit = typename G::Hit;
  const unsigned lane = threadIdx.x & 31;
  const bool thermal = r.letape_th != 0;
  const bool variable_dust = m.p_n_cells != 1;
  if (SM) stage_tables(m, r.p_lambda_in);

  // ---- per-lane packet state ----
  int state = ST_FETCH;
  Rng rng; rng.seed(0, 0, 0);
  double x = 0, y = 0, z = 0, u = 0, v = 0, w = 1;       // interaction / emission point and direction
  double S[4] = {0, 0, 0, 0};
  int lambda = r.lambda_in;
  CellT cell; null_cell(cell);
  bool flag_star = false, flag_scatt = false, flag_ISM = false;
  // flight state (physical_length locals)
  double x0 = 0, y0 = 0, z0 = 0, xo = 0, yo = 0, zo = 0, extr = 0;
  CellT c0, c_old; null_cell(c0); null_cell(c_old);
  DirInv dinv; dinv.inv_a = 0; dinv.inv_w = 0;
  int i_star_hit = 0;
  Rt1Scratch rt1;
  unsigned long long st_steps = 0, st_int = 0, st_sca = 0, st_abs = 0, st_kill = 0, st_esc = 0, st_bounce = 0, st_pk = 0;
  int my_chunk = -1;      // SED-mode chunk bookkeeping

  for (;;) {
    const unsigned need = __ballot_sync(0xffffffffu, state == ST_FETCH);
    if (need == 0 && __all_sync(0xffffffffu, state == ST_DONE)) break;

    // ------------------------------------------------------------ FETCH + EMIT
    if (state == ST_FETCH) {
      // claim the next packet.  Thermal / fixed-count mode: one warp-aggregated atomic on a global
      // counter.  SED mode (dust_transfer.f90:507-510,529,551): each chunk keeps sending until
      // n_photons2 packets were RECEIVED in detector bin capt_sup (or n_phot_lim were sent).
      int nnfot1 = 0;
      unsigned long long idx_in_chunk = 0;
      bool got = false;
      const int first_local = r.nnfot1_start + ((r.rank - ((r.nnfot1_start - 1) % r.n_ranks) + r.n_ranks) % r.n_ranks);
      if (r.count_sent) {
        const unsigned leader = __ffs(need) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(m.work, (unsigned long long)__popc(need));
        base = __shfl_sync(need, base, leader);
        const unsigned long long g = base + __popc(need & ((1u << lane) - 1u));
        if (g < r.n_packets_total) {
          const unsigned long long lc = g / r.n_per_chunk;
          idx_in_chunk = g % r.n_per_chunk;
          nnfot1 = first_local + (int)lc * r.n_ranks;
          my_chunk = (int)lc;
          got = true;
        }
      } else {
        const int start = (int)((blockIdx.x * blockDim.x + threadIdx.x + st_pk) % (unsigned long long)r.n_local_chunks);
        for (int tries = 0; tries < r.n_local_chunks && !got; ++tries) {
          int lc = start + tries; if (lc >= r.n_local_chunks) lc -= r.n_local_chunks;
          if (__ldcg(m.work + 3 + 2 * lc) >= (unsigned long long)r.n_photons2) continue;
          const unsigned long long sidx = atomicAdd(m.work + 2 + 2 * lc, 1ull);
          if (sidx >= r.sent_lim) { atomicAdd(m.work + 2 + 2 * lc, ~0ull); continue; }     // undo (adds -1)
          idx_in_chunk = sidx; nnfot1 = first_local + lc * r.n_ranks; my_chunk = lc; got = true;
        }
      }
      if (!got) state = ST_DONE;
      else {
        rng.seed(r.seed, r.call_index, ((unsigned long long)(nnfot1 - 1) << 40) + idx_in_chunk);
        ++st_pk;
        // n_phot_envoyes(lambda) is incremented with the PREVIOUS packet's lambda in thermal mode
        // (dust_transfer.f90:531 precedes :537); on the device packets are unordered, so the count is
        // attributed to the packet's own emission wavelength (the sum over lambda is identical).
        if (!r.lmono) lambda = select_wl_em<SM>(m, rng.nextf());
        atomicAdd(m.tally + m.lay.n_env + (lambda - 1), 1.0);
        // ---- emit_packet ----
        bool lintersect = true;
        flag_scatt = false;
        float rand = rng.nextf();
        if ((double)rand <= t_frac_star<SM>(m, lambda)) {
          flag_star = true; flag_ISM = false;
          const int i_star = select_star(m, lambda, rng.nextf());
          const float rand1 = rng.nextf(), rand2 = rng.nextf(), rand3 = rng.nextf(), rand4 = rng.nextf();
          // emit_packet_uniform_sphere
          double zz = 2.0 * rand1 - 1.0;
          double srw02 = sqrt(1.0 - zz * zz), sa, ca;
          sincospi(2.0 * rand2 - 1.0, &sa, &ca);            // argmt = pi*(2 rand2 - 1)
          double xx = srw02 * ca, yy = srw02 * sa;
          double cospsi = (double)sqrtf(rand3), sp, cp;
          sincospi(2.0 * (double)rand4, &sp, &cp);          // phi = 2 pi rand4
          cdapres(cospsi, sp, cp, xx, yy, zz, u, v, w);
          const double r_star = m.star[i_star - 1][3] * (1.0 + 1e-6);
          x = xx * r_star + m.star[i_star - 1][0]; y = yy * r_star + m.star[i_star - 1][1]; z = zz * r_star + m.star[i_star - 1][2];
          if (G::is_vor) cell_of_id(m, m.star_icell[i_star - 1], cell);
          else cell = G::index(m, x, y, z);
          if (m.star_out[i_star - 1]) lintersect = G::move_to_grid(m, x, y, z, u, v, w, cell);
          S[0] = m.E_paquet; S[1] = S[2] = S[3] = 0.0;
        } else if ((double)rand <= t_frac_disk<SM>(m, lambda)) {
          flag_star = false; flag_ISM = false;
          const int ic = select_cellule(m, lambda, rng.nextf());
          cell_of_id(m, ic, cell);
          const float rand1 = rng.nextf(), rand2 = rng.nextf(), rand3 = rng.nextf();
          G::pos_em_cell(m, cell, rand1, rand2, rand3, x, y, z);
          random_isotropic_direction(rng, u, v, w);
          S[0] = m.E_paquet; S[1] = S[2] = S[3] = 0.0;
        } else {
          flag_star = false; flag_ISM = true;
          // emit_packet_ISM
          S[0] = 1.0; S[1] = S[2] = S[3] = 0.0;
          const float rand1 = rng.nextf(), rand2 = rng.nextf();
          double zz = 2.0 * rand1 - 1.0;
          double srw02 = sqrt(1.0 - zz * zz), sa, ca;
          sincospi(2.0 * rand2 - 1.0, &sa, &ca);
          double xx = srw02 * ca, yy = srw02 * sa;
          const float rand3 = rng.nextf(), rand4 = rng.nextf();
          double cospsi = (double)(-sqrtf(rand3)), sp, cp;
          sincospi(2.0 * (double)rand4, &sp, &cp);
          cdapres(cospsi, sp, cp, xx, yy, zz, u, v, w);
          x = m.cISM[0] + xx * m.R_ISM; y = m.cISM[1] + yy * m.R_ISM; z = m.cISM[2] + zz * m.R_ISM;
          lintersect = G::move_to_grid(m, x, y, z, u, v, w, cell);
        }
        if (lintersect) state = ST_TAU;
        else {      // packet never enters the model: goes straight to the detector (dust_transfer.f90:545-552)
          if (!flag_ISM) {
            const int capt = capteur(m, r, lambda, u, v, w, S, flag_star, false);
            if (!r.count_sent && capt == r.capt_sup && my_chunk >= 0) atomicAdd(m.work + 3 + 2 * my_chunk, 1ull);
            ++st_esc;
          }
        }
      }
    }

    // ------------------------------------------------------------ TAU: start a flight
    if (state == ST_TAU) {
      const float rand = rng.nextf();
      float tau;
      if (rand == 1.0f) tau = 1.0e30f;
      else if (rand > 1.0e-6f) tau = -logf(1.0f - rand);       // `real` arithmetic in the reference (dust_transfer.f90:1212)
      else tau = rand;
      extr = (double)tau;
      x0 = x; y0 = y; z0 = z; xo = x; yo = y; zo = z;
      c0 = cell; null_cell(c_old);
      dinv = dir_invariants(u, v, w);
      if (!thermal && r.rt1) angles_scatt_rt1(r, u, v, w, rt1);
      i_star_hit = intersect_stars(m, x0, y0, z0, u, v, w);
      state = ST_FLY;
    }

    // ------------------------------------------------------------ FLY: cell crossings
    if (state == ST_FLY) {
#pragma unroll 1
      for (int it = 0; it < 4 && state == ST_FLY; ++it) {
        if (G::test_exit(m, c0, x0, y0, z0)) {
          // the packet leaves the model: detector
          if (!flag_ISM) {
            const int capt = capteur(m, r, lambda, u, v, w, S, flag_star, flag_scatt);
            if (!r.count_sent && capt == r.capt_sup && my_chunk >= 0) atomicAdd(m.work + 3 + 2 * my_chunk, 1ull);
          }
          ++st_esc;
          state = ST_FETCH;
          break;
        }
        if (i_star_hit > 0) {
          CellT cs; cell_of_id(m, m.star_icell[i_star_hit - 1], cs);
          if (same_cell(c0, cs)) { ++st_kill; state = ST_FETCH; break; }     // packet absorbed by the star
        }
        const int idx = tally_index(m, c0);
        double opacity = 0.0;
        int p_icell = 1;
        if (idx >= 0) {
          p_icell = variable_dust ? idx + 1 : 1;
          opacity = t_kappa<SM>(m, p_icell, lambda) * __ldg(m.kappa_factor + idx);
          if (__ldg(m.dark + idx)) {
            // dark-zone bounce (optical_depth.f90:104-112): back to the previous cell's entry point, reversed
            u = -u; v = -v; w = -w;
            cell = c_old; x = xo; y = yo; z = zo;
            ++st_bounce;
            state = ST_INTERACT;
            break;
          }
        }
        const Hit h = G::distance(m, dinv, x0, y0, z0, u, v, w, c0, c_old);
        ++st_steps;
        double l_contrib = hit_l_contrib(h), l = h.l;
        const double tau_c = l_contrib * opacity;
        bool lstop = false;
        if (tau_c > extr) {
          lstop = true;
          l_contrib = l_contrib * (extr / tau_c);
          l = hit_l_void(h) + l_contrib;
        } else extr = extr - tau_c;
        if (idx >= 0) {
          // save_radiation_field
          if (thermal) {
            atomicAdd(m.tally + m.lay.xKJ + idx, t_kappa_abs<SM>(m, p_icell, lambda) * l_contrib * S[0]);
            if (r.lxJ) atomicAdd(m.tally + m.lay.xJ + idx + (size_t)m.n_cells * (lambda - 1), l_contrib * S[0]);
          } else {
            if (r.lxJ) atomicAdd(m.tally + m.lay.xJ + idx + (size_t)m.n_cells * (lambda - 1), l_contrib * S[0]);
            if (r.rt1) {
              double x1, y1, z1;
              G::exit_point(h, x0, y0, z0, u, v, w, x1, y1, z1);
              deposit_rt1(m, r, idx, p_icell, r.p_lambda_in, l_contrib, S, flag_star,
                          0.5 * (x0 + x1), 0.5 * (y0 + y1), 0.5 * (z0 + z1), rt1);
            }
          }
        }
        if (lstop) {
          x = x0 + l * u; y = y0 + l * v; z = z0 + l * w;
          cell = c0;
          if (!G::is_vor && m.l3D && m.kind == 1) cell = G::index(m, x, y, z);     // optical_depth.f90:162-165
          state = ST_INTERACT;
        } else {
          double x1, y1, z1;
          CellT c1;
          G::advance(m, h, x0, y0, z0, u, v, w, c0, x1, y1, z1, c1);
          xo = x0; yo = y0; zo = z0; c_old = c0;
          x0 = x1; y0 = y1; z0 = z1; c0 = c1;
        }
      }
    }

    // ------------------------------------------------------------ INTERACT
    if (state == ST_INTERACT) {
      ++st_int;
      const int idx = tally_index(m, cell);
      const int p_icell = (variable_dust && idx >= 0) ? idx + 1 : 1;
      const float albedo = t_albedo<SM>(m, p_icell, lambda);
      float rand;
      bool dead = false;
      if (r.lmono) {      // forced scattering (dust_transfer.f90:1263-1278)
        if (idx >= 0 && __ldg(m.dark + idx)) dead = true;
        else {
          S[0] *= albedo; S[1] *= albedo; S[2] *= albedo; S[3] *= albedo;
          if (S[0] < (double)(FLT_MIN * 1.0e6f)) dead = true;
        }
        rand = -1.0f;
      } else rand = rng.nextf();
      if (dead) { ++st_kill; state = ST_FETCH; }
      else if (rand < albedo) {
        // ---- scattering, method 2 (dust_transfer.f90:1318-1348)
        ++st_sca;
        flag_scatt = true;
        rand = rng.nextf();
        const float rand2 = rng.nextf();
        int itheta; double cospsi;
        if (r.lmethod_aniso1) angle_diff_theta_pos<SM>(m, r.p_lambda_in, p_icell, rand, rand2, itheta, cospsi);
        else hg(t_gfac<SM>(m, p_icell, lambda), rand, itheta, cospsi);
        if (r.lisotropic) { itheta = 1; cospsi = (double)__fsub_rn(__fmul_rn(2.0f, rand), 1.0f); }
        rand = rng.nextf();
        double sp, cp;
        sincospi((double)__fsub_rn(__fmul_rn(2.0f, rand), 1.0f), &sp, &cp);     // PHI = PI*(2.0*rand-1.0): fp32 inner
        double u1, v1, w1;
        cdapres(cospsi, sp, cp, u, v, w, u1, v1, w1);
        if (r.lmethod_aniso1 && r.lsepar_pola) scatter_stokes(m, lambda, itheta, rand2, p_icell, S, u, v, w, u1, v1, w1);
        u = u1; v = v1; w = w1;
        state = ST_TAU;
      } else {
        // ---- absorption + immediate re-emission (LTE) (dust_transfer.f90:1353-1402)
        ++st_abs;
        flag_star = false; flag_scatt = false; flag_ISM = false;
        (void)rng.nextf();                       // rand1 is drawn but unused in the high-memory LTE branch
        const float rand2 = rng.nextf();
        lambda = im_reemission_LTE<SM>(m, r, idx, p_icell, rand2);
        random_isotropic_direction(rng, u, v, w);
        S[1] = 0.0; S[2] = 0.0; S[3] = 0.0;
        state = ST_TAU;
      }
    }
  }

  // ---- diagnostics (not part of the reference) ----
  double* st = m.tally + m.lay.stats;
  auto flush = [&](int k, unsigned long long vv) {
    for (int o = 16; o > 0; o >>= 1) vv += __shfl_down_sync(0xffffffffu, vv, o);
    if (lane == 0 && vv) atomicAdd(st + k, (double)vv);
  };
  flush(STAT_PACKETS, st_pk); flush(STAT_STEPS, st_steps); flush(STAT_INTERACT, st_int); flush(STAT_SCATT, st_sca);
  flush(STAT_ABS, st_abs); flush(STAT_KILLED, st_kill); flush(STAT_ESCAPED, st_esc); flush(STAT_BOUNCE, st_bounce);
}

}  // namespace mcb
