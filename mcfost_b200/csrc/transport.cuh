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
// Execution model (B200): persistent blocks, one per SM, each owning a pool of
// NP = 1024 packets whose state lives in SHARED MEMORY (structure of arrays,
// ~124 B per packet).  Work is organised in rounds: every packet sits in exactly
// one per-phase queue (EMIT, FLY, SCATTER, ABSORB); warps claim 32-entry chunks
// of a single queue, run that phase for the 32 packets (registers <-> shared
// memory), and push each packet to the queue of its next phase with one warp-
// aggregated shared-memory atomic.  This is the "periodic regrouping": every
// warp instruction is executed by lanes that are all in the same phase, and the
// RNG state of a packet is just (packet id, event counter), so any thread can
// continue any packet.  New packets are claimed from a global counter with one
// warp-aggregated atomic.  The small hot tables (walls, kappa(lambda), albedo,
// log Qcool, the k dB/dT CDF, the s11 CDF, cos table, emission spectra; ~48 KB
// for ref4.1) are staged once per block in shared memory when the dust is not
// cell-dependent; per-cell arrays stay in global memory behind L1/L2.  Tallies
// are L2 atomics (red.global.add.f64); the running cell temperature reads them
// back with ld.global.cg (L1 is not coherent with L2 atomics).  The next-cell
// half of a crossing is skipped when the flight ends inside the cell.
#pragma once
#include "model.cuh"
#include "philox.cuh"
#include "cells.cuh"

namespace mcb {

// The model and run parameters of the launch in flight live in constant memory (written by
// mcb_launch_mc on the handle's stream): every phase function reads them through the constant
// bank without threading pointers through the non-inlined calls.  There are MCB_BANKS copies so
// that launches of up to three handles can be in flight together (the drain-out of one call overlaps
// the next calls of the other handles); mc_kernel.cu guards a bank against a second concurrent user.
constexpr int MCB_BANKS = 3;
__constant__ DevModel c_mm[MCB_BANKS];
__constant__ DevRun c_rr[MCB_BANKS];
// BANK is a template parameter of every function below that touches the constants, so that
// c_m / c_r are compile-time constant-bank addresses (no indexed constant loads in the hot path)
#define c_m (c_mm[BANK])
#define c_r (c_rr[BANK])

// kernel variants (template parameter VAR): the default thermal step gets its own instantiation without the SED /
// image-step code (forced scattering, received-packet chunk accounting, ray-tracing accumulators) and without the
// rarely used options; every such branch that merely sat behind a run-time flag cost instruction-cache footprint
// and registers of the hot loop.
enum { VAR_THERMAL = 0, VAR_GENERIC = 1, VAR_EXTRAS = 2 };
enum { Q_EMIT = 0, Q_ABS = 1, Q_SCAT = 2, Q_FLY = 3, NQ = 4, Q_NONE = 7 };     // queue order = claim order
enum { CTL_LIVE = 2 * NQ, CTL_BUSY, CTL_PARK, CTL_DRY, CTL_SENT };      // Pool::ctl: [0, NQ) heads, [NQ, 2 NQ) tails, then these
enum { STAT_PACKETS = 0, STAT_STEPS, STAT_INTERACT, STAT_SCATT, STAT_ABS, STAT_KILLED, STAT_ESCAPED, STAT_BOUNCE, STAT_MRW_WALKS, STAT_MRW_STEPS };

// ---- opacity / thermal table accessors (SM: shared-memory staging, p_n_cells == 1) ----
template <bool SM> __device__ __forceinline__ double t_kappa(const DevModel& m, int p_icell, int lambda) {
  return SM ? smd_ld(m, m.sm.kappa + lambda - 1) : __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ double t_kappa_abs(const DevModel& m, int p_icell, int lambda) {
  return SM ? smd_ld(m, m.sm.kappa_abs + lambda - 1) : __ldg(m.kappa_abs + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ float t_albedo(const DevModel& m, int p_icell, int lambda) {
  return SM ? smf_ld(m, m.sm.albedo, lambda - 1) : __ldg(m.albedo + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ float t_gfac(const DevModel& m, int p_icell, int lambda) {
  return SM ? smf_ld(m, m.sm.gfac, lambda - 1) : __ldg(m.gfac + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)); }
template <bool SM> __device__ __forceinline__ double t_logQ(const DevModel& m, int t, int p_icell) {      // t 1-based
  return SM ? smd_ld(m, m.sm.logQ + t - 1) : __ldg(m.logQ + (size_t)m.n_T * (p_icell - 1) + t - 1); }
template <bool SM> __device__ __forceinline__ double t_kdB(const DevModel& m, int l, int t, int p_icell) {  // l, t 1-based
  return (SM && m.sm.kdB >= 0) ? smd_ld(m, m.sm.kdB + (l - 1) + m.n_lambda * (t - 1))
            : __ldg(m.kdB + (size_t)m.n_lambda * ((t - 1) + (size_t)m.n_T * (p_icell - 1)) + (l - 1)); }
template <bool SM> __device__ __forceinline__ double t_cos(const DevModel& m, int k) { return SM ? smd_ld(m, m.sm.cos_tab + k) : __ldg(m.cos_tab + k); }
template <bool SM> __device__ __forceinline__ float t_prob_s11(const DevModel& m, int k, int p_icell, int p_lambda) {
  return SM ? smf_ld(m, m.sm.prob_s11, k) : __ldg(m.prob_s11 + (size_t)(NANG + 1) * ((p_icell - 1) + (size_t)m.p_n_cells * (p_lambda - 1)) + k); }
template <bool SM> __device__ __forceinline__ double t_spec_cumul(const DevModel& m, int k) { return SM ? smd_ld(m, m.sm.spec_cumul + k) : __ldg(m.spec_cumul + k); }
template <bool SM> __device__ __forceinline__ double t_frac_star(const DevModel& m, int lambda) { return SM ? smd_ld(m, m.sm.frac_star + lambda - 1) : __ldg(m.frac_star + lambda - 1); }
template <bool SM> __device__ __forceinline__ double t_frac_disk(const DevModel& m, int lambda) { return SM ? smd_ld(m, m.sm.frac_disk + lambda - 1) : __ldg(m.frac_disk + lambda - 1); }

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
  cp(L.logQ, m.logQ, m.n_T); if (L.kdB >= 0) cp(L.kdB, m.kdB, m.n_lambda * m.n_T);
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
    const double q = 1.0 - w0 * w0, cm1 = rsqrt(q), c = q * cm1, aw0 = a * w0;      // (sqrt and its reciprocal from one rsqrt)
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
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- heavy libm entry points, not inlined: one copy each in the instruction stream ----
__device__ __noinline__ double mcb_log(double x) { return log(x); }
// log of a positive normal number to ~2e-7 absolute: exponent * ln 2 + MUFU.LG2 of the mantissa.  Used where the result is
// compared with a 100-point cooling table (Temp_LTE: 2e-7 in log Q is 2e-7 / 0.3 of a temperature bin, a relative error of
// 1e-7 on T) -- ~12 instructions instead of the ~60 of the double-precision log on the latency-critical chain.
__device__ __forceinline__ double mc_log_table(double x) {
  const long long b = __double_as_longlong(x);
  const int e = (int)((b >> 52) & 0x7ff) - 1023;
  const double mant = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);      // [1, 2)
  return fma((double)e, 0.6931471805599453, (double)(__log2f((float)mant) * 0.69314718f));
}
__device__ __noinline__ void mcb_sincospi(double x, double* s, double* c) { sincospi(x, s, c); }

// ---- random_numbers.f90:32-51 (the two draws are passed in) -------------------
__device__ __forceinline__ void random_isotropic_direction(float rand_w, float rand_phi, double& u, double& v, double& w) {
  w = 2.0 * rand_w - 1.0;
  double uv = sqrt(1.0 - w * w);
  double sp, cp;
  mcb_sincospi(2.0 * rand_phi - 1.0, &sp, &cp);       // phi = pi*(2 rand - 1)
  u = uv * cp; v = uv * sp;
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
    double q = mc_div(1.0 - g2, 1.0 - g1 + 2.0 * g1 * rand_dp);
    cospsi = mc_div(1.0 + g2 - q * q, 2.0 * g1);
  } else cospsi = 2.0 * rand_dp - 1.0;
  itheta = (int)floor(acos(cospsi) * (180.0 / MCB_PI)) + 1;
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

// ---- scattering.f90:1187-1298 update_Stokes with the per-cell Mueller matrix of
// get_Mueller_matrix_per_cell (:1328-1350); sparse products written out.
// The reference builds the rotation from angles (theta = atan2(v1,u1) in `rotation`, utils.f90:584;
// theta = acos(costhet), omega = 2(theta + pi/2), cos/sin(omega) in fp32, scattering.f90:1242-1261).
// Here the same quantities are formed algebraically (cos(atan2(v,u)) = u/hypot(u,v);
// cos(omega) = 1 - 2 costhet^2, sin(omega) = -+2 costhet sqrt(1 - costhet^2)): identical up to the
// fp32 rounding the reference itself carries, without atan2 / sincos / acosf / cosf / sinf.
__device__ __forceinline__ void stokes_update(double M11, double M12, double M22, double M33, double M34, double M44, double* S,
                                              double u0, double v0, double w0, double u1, double v1, double w1) {
  const double M43 = -M34;
  // rotation(u0,v0,w0 ; u1,v1,w1) -> v1p (utils.f90:553-599), algebraic cos/sin of atan2(v1,u1)
  double cost, sint, sing;
  if (w1 > 0.999999999) { cost = 1.0; sint = 0.0; sing = 0.0; }
  else if (fabs(u1) < MCB_TINY_REAL) { cost = 0.0; sint = 1.0; sing = sqrt(1.0 - w1 * w1); }
  else { const double ih = rsqrt(u1 * u1 + v1 * v1); cost = u1 * ih; sint = v1 * ih; sing = sqrt(1.0 - w1 * w1); }
  const double prod = cost * u0 + sint * v0;
  const double v1pj = cost * v0 - sint * u0;
  const double v1pk = sing * w0 - w1 * prod;
  const float xnyp = (float)sqrt(v1pk * v1pk + v1pj * v1pj);
  float costhet;
  if (xnyp < 1e-10f) costhet = 1.0f;
  else costhet = (float)mc_div(-1.0 * v1pj, (double)xnyp);
  costhet = fminf(1.0f, fmaxf(-1.0f, costhet));
  float cosw = 1.0f - 2.0f * costhet * costhet;                          // cos(2 theta + pi)
  float sinw = -2.0f * costhet * sqrtf(fmaxf(0.0f, 1.0f - costhet * costhet));   // sin(2 theta + pi), theta in [0, pi]
  if (v1pk < 0.0) sinw = -sinw;
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
    const double f = mc_div(M11 * S1_0, S[0]);
    S[0] *= f; S[1] *= f; S[2] *= f; S[3] *= f;
  }
}
// per-cell Mueller matrix (get_Mueller_matrix_per_cell, scattering.f90:1328-1350): M11 = 1
template <int BANK>
__device__ __forceinline__ void scatter_stokes(int lambda, int itheta, float frac, int p_icell, double* S,
                                               double u0, double v0, double w0, double u1, double v1, double w1) {
  const DevModel& m = c_m;
  const size_t q1 = (size_t)itheta + (size_t)(NANG + 1) * ((p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)), q0 = q1 - 1;
  const float frac_m1 = 1.0f - frac;
  const float a22 = __ldg(m.s22 + q1), b22 = __ldg(m.s22 + q0), a12 = __ldg(m.s12 + q1), b12 = __ldg(m.s12 + q0);
  const float a33 = __ldg(m.s33 + q1), b33 = __ldg(m.s33 + q0), a44 = __ldg(m.s44 + q1), b44 = __ldg(m.s44 + q0);
  const float a34 = __ldg(m.s34 + q1), b34 = __ldg(m.s34 + q0);
  const double M11 = (double)1.0f;
  const double M22 = (double)__fadd_rn(__fmul_rn(a22, frac), __fmul_rn(b22, frac_m1));
  const double M12 = (double)__fadd_rn(__fmul_rn(a12, frac), __fmul_rn(b12, frac_m1));
  const double M33 = (double)__fadd_rn(__fmul_rn(a33, frac), __fmul_rn(b33, frac_m1));
  const double M44 = (double)__fadd_rn(__fmul_rn(a44, frac), __fmul_rn(b44, frac_m1));
  const double M34 = (double)__fsub_rn(__fmul_rn(-a34, frac), __fmul_rn(b34, frac_m1));
  stokes_update(M11, M12, M22, M33, M34, M44, S, u0, v0, w0, u1, v1, w1);
}

// ---- thermal_emission.f90:649-771 Temp_LTE + im_reemission_LTE (high-memory
// branch): new wavelength index ------------------------------------------------
// The three global loads it starts from (running tally, cached temperature index, cell volume) are passed in
// so that the caller can issue them BEFORE the Philox blocks of the event: one L2 round trip of the dependent
// chain instead of three (matters for the last, latency-bound packets of a call).
struct LtePre { double xkj, vol; int Ti; };
__device__ __forceinline__ LtePre lte_prefetch(const DevModel& m, int idx) {
  LtePre p;
  p.xkj = __ldcg(m.tally + m.lay.xKJ + idx);      // running tally: L2-coherent load (the adds are L2 atomics)
  p.Ti = __ldcg(m.xT_ech + idx);
  p.vol = __ldg(m.volume + idx);
  return p;
}
// Temp_LTE (thermal_emission.f90:649-706): temperature index and interpolation fraction of the cell
template <bool SM>
__device__ __forceinline__ void temp_lte(const DevModel& m, const DevRun& r, int idx, int p_icell, const LtePre pre, int& Ti_out, double& frac_out) {
  double Qheat = mc_div(pre.xkj * r.nb_proc_equiv * m.L_packet_th, pre.vol);
  int Ti = 2;
  double frac_T2 = 0.0;       // `frac` is left undefined by the reference at T_min; 0 chosen (same as the oracle)
  if (!(Qheat < MCB_TINY_DP)) {
    double log_Qheat = mc_log_table(Qheat);
    if (!(log_Qheat < t_logQ<SM>(m, 1, p_icell))) {
      Ti = pre.Ti;
      while ((t_logQ<SM>(m, Ti, p_icell) < log_Qheat) && (Ti < m.n_T)) ++Ti;
      // another warp may have cached an index computed from a larger running tally: step back down
      while (Ti > 2 && !(t_logQ<SM>(m, Ti - 1, p_icell) < log_Qheat)) --Ti;
      double q1 = t_logQ<SM>(m, Ti - 1, p_icell), q2 = t_logQ<SM>(m, Ti, p_icell);
      frac_T2 = mc_div(log_Qheat - q1, q2 - q1);
    }
  }
  if (Ti > pre.Ti) atomicMax(m.xT_ech + idx, Ti);      // (the cache only grows: nothing to write for a packet that finds it current)
  Ti_out = Ti; frac_out = frac_T2;
}
template <bool SM>
__device__ __forceinline__ int im_reemission_LTE(const DevModel& m, const DevRun& r, int idx, int p_icell, float rand2, const LtePre pre) {
  int Ti; double frac_T2;
  temp_lte<SM>(m, r, idx, p_icell, pre, Ti, frac_T2);
  const double frac_T1 = 1.0 - frac_T2;
  int l1 = 0, l2 = m.n_lambda, l = (l1 + l2) / 2;
  while ((l2 - l1) > 1) {
    double proba = frac_T1 * t_kdB<SM>(m, l, Ti - 1, p_icell) + frac_T2 * t_kdB<SM>(m, l, Ti, p_icell);
    if ((double)rand2 > proba) l1 = l; else l2 = l;
    l = (l1 + l2) / 2;
  }
  return l + 1;
}


// =============================================================================
// Per-grain modes: scattering method 1 (dust_transfer.f90:1291-1317) and the nLTE / qRE
// re-emission branches (dust_transfer.f90:1353-1395).  Cold code: __noinline__, global-memory tables, and
// compiled only into the GR = true instantiation of the kernel, so the LTE / method-2 kernels are unchanged.
// GR = true also carries the other rarely used options (complete capteur with photon maps / origin tallies, hot
// spot, weighted emission, low-memory LTE emission, xN_abs): each of them cost steady-state throughput of the
// default kernel when it merely sat behind a run-time flag (measured -10 % for all of them together).
// =============================================================================
#define MCB_AU_TO_CM_MUM2 ((149597870700.0 * 100.0) * (1.0e-4 * 1.0e-4))     /* AU_to_cm * mum_to_cm**2 */

// thermal_emission.f90:1953-2040 select_absorbing_grain, heating_method 1 (LTE) / 2 (nLTE) / 3 (qRE); idx 0-based cell
template <int BANK>
__device__ __noinline__ int select_absorbing_grain(int lambda, int idx, float rand, int heating_method) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  const bool variable = m.p_n_cells != 1;
  const size_t pl = (size_t)(variable ? idx : 0) + (size_t)m.p_n_cells * (lambda - 1);
  const double kf = __ldg(m.kappa_factor + idx);
  double norm; int kstart, kend;
  if (heating_method == 1) {
    norm = __ldg(m.kappa_abs + pl) * kf / MCB_AU_TO_CM_MUM2;
    kstart = g.LTE_s; kend = g.LTE_e;
  } else if (heating_method == 2) {
    norm = __ldg(g.kappa_abs_nLTE + pl) * kf / MCB_AU_TO_CM_MUM2;
    kstart = g.nLTE_s; kend = g.nLTE_e;
  } else {
    const double kRE = __ldg(g.kappa_abs_RE + idx + (size_t)m.n_cells * (lambda - 1));
    if (r.lRE_nLTE) norm = (kRE - (__ldg(m.kappa_abs + pl) + __ldg(g.kappa_abs_nLTE + pl)) * kf) / MCB_AU_TO_CM_MUM2;
    else            norm = (kRE - __ldg(m.kappa_abs + pl) * kf) / MCB_AU_TO_CM_MUM2;
    kstart = g.nRE_s; kend = g.nRE_e;
  }
  const bool masked = heating_method > 2;
  const int nkR = g.nRE_e - g.nRE_s + 1;
  const bool up = rand < 0.5f;
  const double prob = up ? rand * norm : (double)(1.0f - rand) * norm;
  double CDF = 0.0;
  int k = up ? kstart : kend;
  const int step = up ? 1 : -1;
  for (; up ? (k <= kend) : (k >= kstart); k += step) {
    if (!masked || __ldg(g.l_RE + (k - g.nRE_s) + (size_t)nkR * idx))
      CDF = CDF + (double)__ldg(g.C_abs + (k - 1) + (size_t)g.n_grains_tot * (lambda - 1)) * __ldg(g.dd + (variable ? k - 1 : 0) + (size_t)g.n_dens * idx) * __ldg(g.n_grains + k - 1);
    if (CDF > prob) break;
  }
  // the reference returns the run-out do-variable (kend+1 / kstart-1) and then indexes out of bounds
  // (:2033-2034 are commented out); only reachable with inconsistent tables: clamp for memory safety
  return min(max(k, kstart), kend);
}

// shared tail of im_reemission_NLTE / im_reemission_qRE (thermal_emission.f90:812-863, 1469-1511):
// temperature index of grain kk (0-based inside its regime) from log_E_abs, then the wavelength bisection
template <int BANK>
__device__ __noinline__ int reemit_1grain(int* xT, const double* logE, const double* kdB, int nk, int kk, int idx,
                                          double log_E_abs, float rand2) {
  const DevModel& m = c_m;
  int* px = xT + kk + (size_t)nk * idx;
  auto LE = [&](int t) { return __ldg(logE + kk + (size_t)nk * (t - 1)); };
  int Ti = __ldcg(px);
  while ((LE(Ti) < log_E_abs) && (Ti < m.n_T)) ++Ti;
  while (Ti > 2 && !(LE(Ti - 1) < log_E_abs)) --Ti;      // index cached by another warp from a later, larger tally
  atomicMax(px, Ti);
  const int T2 = Ti, T1 = Ti - 1;
  const double Temp2 = __ldg(m.tab_Temp + T2 - 1), Temp1 = __ldg(m.tab_Temp + T1 - 1);
  const double frac = (log_E_abs - LE(T1)) / (LE(T2) - LE(T1));
  const double Temp = exp(mcb_log(Temp2) * frac + mcb_log(Temp1) * (1.0 - frac));
  const double frac_T2 = (Temp - Temp1) / (Temp2 - Temp1), frac_T1 = 1.0 - frac_T2;
  auto CDF = [&](int l, int t) { return __ldg(kdB + (l - 1) + (size_t)m.n_lambda * (kk + (size_t)nk * (t - 1))); };
  int l1 = 0, l2 = m.n_lambda, l = (l1 + l2) / 2;
  while ((l2 - l1) > 1) {
    const double proba = frac_T1 * CDF(l, T1) + frac_T2 * CDF(l, T2);
    if ((double)rand2 > proba) l1 = l; else l2 = l;
    l = (l1 + l2) / 2;
  }
  return l + 1;
}

// Sum_lambda C_abs_norm(k,l) * (xJ_abs(icell,l) [running, all ranks] + J0(icell, l or lambda0))
template <int BANK>
__device__ __forceinline__ double log_E_abs_1grain(int k, int idx, int j0_fixed_lambda) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  double J_abs = 0.0;
  for (int il = 1; il <= m.n_lambda; ++il) {
    const size_t cl = (size_t)idx + (size_t)m.n_cells * (il - 1);
    const double j0 = __ldg(g.J0 + (j0_fixed_lambda ? (size_t)idx + (size_t)m.n_cells * (j0_fixed_lambda - 1) : cl));
    J_abs = J_abs + (double)__ldg(g.C_abs_norm + (k - 1) + (size_t)g.n_grains_tot * (il - 1)) * (__ldcg(m.tally + m.lay.xJ + cl) * r.nb_proc_equiv + j0);
  }
  return mcb_log(J_abs * m.L_packet_th / __ldg(m.volume + idx));
}

// thermal_emission.f90:775-866 im_reemission_NLTE
template <int BANK>
__device__ __noinline__ int im_reemission_NLTE(int idx, int lambda0, float rand1, float rand2) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  int k;
  if (r.low_mem_nLTE) k = select_absorbing_grain<BANK>(lambda0, idx, rand1, 2);
  else {
    const int nk1 = g.nLTE_e - g.nLTE_s + 2;
    const double* cdf = g.kabs_nLTE_CDF + (size_t)nk1 * ((size_t)idx + (size_t)m.n_cells * (lambda0 - 1)) - (g.nLTE_s - 1);
    int kmin = g.nLTE_s, kmax = g.nLTE_e;
    k = (kmin + kmax) / 2;
    while ((kmax - kmin) > 1) {
      if (__ldg(cdf + k) < (double)rand1) kmin = k; else kmax = k;
      k = (kmin + kmax) / 2;
    }
    k = kmax;
  }
  const double log_E_abs = log_E_abs_1grain<BANK>(k, idx, 0);
  return reemit_1grain<BANK>(g.xT_1g, g.logE, g.kdB, g.nLTE_e - g.nLTE_s + 1, k - g.nLTE_s, idx, log_E_abs, rand2);
}
// thermal_emission.f90:1441-1514 im_reemission_qRE (J0(icell,lambda) with the ABSORBED wavelength, :1466)
template <int BANK>
__device__ __noinline__ int im_reemission_qRE(int idx, int lambda0, float rand1, float rand2) {
  const DevModel& m = c_m; const DevGrains& g = m.gr;
  const int k = select_absorbing_grain<BANK>(lambda0, idx, rand1, 3);
  const double log_E_abs = log_E_abs_1grain<BANK>(k, idx, lambda0);
  return reemit_1grain<BANK>(g.xT_1g_nRE, g.logE_nRE, g.kdB_nRE, g.nRE_e - g.nRE_s + 1, k - g.nRE_s, idx, log_E_abs, rand2);
}

// dust_prop.f90:1292-1380 select_scattering_grain (the reference passes p_icell as its `icell`)
template <int BANK>
__device__ __noinline__ int select_scattering_grain(int lambda, int p_icell, float rand) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  const bool variable = m.p_n_cells != 1;
  const size_t pl = (size_t)(p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1);
  if (r.low_mem_scattering) {
    const double norm = __ldg(m.kappa + pl) * __ldg(m.albedo + pl) / MCB_AU_TO_CM_MUM2;
    const bool up = rand < 0.5f;
    const double prob = up ? rand * norm : (double)(1.0f - rand) * norm;
    double CDF = 0.0;
    int k = up ? 1 : g.n_grains_tot;
    const int step = up ? 1 : -1;
    for (; up ? (k <= g.n_grains_tot) : (k >= 1); k += step) {
      const int pk = variable ? k : __ldg(g.zone + k - 1);
      const double density = __ldg(g.dd + (pk - 1) + (size_t)g.n_dens * (p_icell - 1)) * __ldg(g.n_grains + k - 1);
      CDF = CDF + __ldg(g.C_sca + (k - 1) + (size_t)g.n_grains_tot * (lambda - 1)) * density;
      if (CDF > prob) break;
    }
    return min(max(k, 1), g.n_grains_tot);
  }
  const double* cdf = g.ksca_CDF + (size_t)(g.n_grains_tot + 1) * pl;
  const double prob = (double)rand;
  int kmin = 0, kmax = g.n_grains_tot, k = (kmin + kmax) / 2;
  while (__ldg(cdf + k) != prob) {
    if (__ldg(cdf + k) < prob) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
    if ((kmax - kmin) <= 1) break;
  }
  return kmax;
}
// scattering.f90:1387-1429 angle_diff_theta (per grain)
template <int BANK>
__device__ __noinline__ void angle_diff_theta_grain(int lambda, int igrain, float rand, float rand2, int& itheta, double& cospsi) {
  const DevModel& m = c_m; const DevGrains& g = m.gr;
  const float* p = g.prob_s11 + (lambda - 1) + (size_t)m.n_lambda * (igrain - 1);
  const size_t stride = (size_t)m.n_lambda * g.n_grains_tot;
  int kmin = 0, kmax = NANG, k = (kmin + kmax) / 2;
  while ((kmax - kmin) > 1) {
    if (__ldg(p + stride * k) < rand) kmin = k; else kmax = k;
    k = (kmin + kmax) / 2;
  }
  k = kmax;
  itheta = k;
  const double c0 = __ldg(m.cos_tab + k - 1), c1 = __ldg(m.cos_tab + k);
  cospsi = c0 + rand2 * (c1 - c0);
}


// dust_transfer.f90:1291-1317: scattering method 1.  Philox words of the interaction block:
// x = grain draw, y = rand, z = rand2, w = phi draw.
// (results come back by value: no local of the hot SCATTER / ABSORB phases has its address taken)
struct Scat1Out { double u1, v1, w1, S0, S1, S2, S3; };
template <int BANK>
__device__ __noinline__ Scat1Out scatter_method1(int lambda, int p_icell, uint4 b, bool pola, double S0, double S1, double S2, double S3,
                                                 double u, double v, double w) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  double S[4] = {S0, S1, S2, S3};
  double u1, v1, w1;
  const int igrain = select_scattering_grain<BANK>(lambda, p_icell, u01(b.x));
  const float rand = u01(b.y), rand2 = u01(b.z), rand3 = u01(b.w);
  int itheta; double cospsi;
  if (r.lmethod_aniso1) angle_diff_theta_grain<BANK>(lambda, igrain, rand, rand2, itheta, cospsi);
  else {
    hg(__ldg(g.tab_g + (igrain - 1) + (size_t)g.n_grains_tot * (lambda - 1)), rand, itheta, cospsi);
    if (r.lisotropic) { itheta = 1; cospsi = (double)__fsub_rn(__fmul_rn(2.0f, rand), 1.0f); }
  }
  double sp, cp;
  mcb_sincospi((double)__fsub_rn(__fmul_rn(2.0f, rand3), 1.0f), &sp, &cp);
  cdapres(cospsi, sp, cp, u, v, w, u1, v1, w1);
  if (pola && r.lmethod_aniso1) {     // get_Mueller_matrix_per_grain, scattering.f90:1302-1324 (fp32 interpolation)
    const size_t q1 = (size_t)itheta + (size_t)(NANG + 1) * ((igrain - 1) + (size_t)g.n_grains_tot * (lambda - 1)), q0 = q1 - 1;
    const float frac = rand2, frac_m1 = 1.0f - frac;
    auto mix = [&](const float* t) { return (double)__fadd_rn(__fmul_rn(__ldg(t + q1), frac), __fmul_rn(__ldg(t + q0), frac_m1)); };
    const double M34 = (double)__fsub_rn(__fmul_rn(-__ldg(g.s34 + q1), frac), __fmul_rn(__ldg(g.s34 + q0), frac_m1));
    stokes_update(mix(g.s11), mix(g.s12), mix(g.s22), mix(g.s33), M34, mix(g.s44), S, u, v, w, u1, v1, w1);
  }
  return Scat1Out{u1, v1, w1, S[0], S[1], S[2], S[3]};
}

// im_reemission_LTE, low-memory branch (thermal_emission.f90:739-751): the absorbing LTE grain is drawn and its own
// kdB_dT_1grain_LTE_CDF is bisected instead of the cell's kdB_dT_CDF
template <bool SM, int BANK>
__device__ __noinline__ int im_reemission_LTE_lowmem(int idx, int p_icell, int lambda0, float rand1, float rand2, const LtePre pre) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  int Ti; double frac_T2;
  temp_lte<SM>(m, r, idx, p_icell, pre, Ti, frac_T2);
  const double frac_T1 = 1.0 - frac_T2;
  const int k = select_absorbing_grain<BANK>(lambda0, idx, rand1, 1);
  const int nk = g.LTE_e - g.LTE_s + 1, kk = k - g.LTE_s;
  auto CDF = [&](int l, int t) { return __ldg(g.kdB_LTE + (l - 1) + (size_t)m.n_lambda * (kk + (size_t)nk * (t - 1))); };
  int l1 = 0, l2 = m.n_lambda, l = (l1 + l2) / 2;
  while ((l2 - l1) > 1) {
    const double proba = frac_T1 * CDF(l, Ti - 1) + frac_T2 * CDF(l, Ti);
    if ((double)rand2 > proba) l1 = l; else l2 = l;
    l = (l1 + l2) / 2;
  }
  return l + 1;
}

// dust_transfer.f90:1353-1395 when .not.lonly_LTE: energy kept by grains out of equilibrium, choice of
// the grain regime, re-emission wavelength.  Returns the new wavelength, or 0 if the packet is dropped.
// S0 is scaled in place; e_nRE returns this packet's contribution to E_abs_nRE.
struct AbsOut { int lambda; double S0, e_nRE; };
template <bool SM, int BANK>
__device__ __noinline__ AbsOut absorb_grain_regimes(int idx, int p_icell, int lambda0, uint4 b, float sel, double S0) {
  const DevModel& m = c_m; const DevRun& r = c_r; const DevGrains& g = m.gr;
  const size_t cl = (size_t)idx + (size_t)m.n_cells * (lambda0 - 1);
  AbsOut o = {0, S0, 0.0};
  if (r.lnRE) {
    const double pRE = __ldg(g.proba_abs_RE + cl);
    o.e_nRE = S0 * (1.0 - pRE);
    o.S0 = S0 * pRE;
    if (o.S0 < MCB_TINY_REAL) return o;
  }
  const float rand1 = u01(b.x), rand2 = u01(b.y);
  if (r.lonly_nLTE) o.lambda = im_reemission_NLTE<BANK>(idx, lambda0, rand1, rand2);
  else if ((double)sel <= __ldg(g.P_LTE + cl)) o.lambda = r.low_mem_th ? im_reemission_LTE_lowmem<SM, BANK>(idx, p_icell, lambda0, rand1, rand2, lte_prefetch(m, idx))
                                                                     : im_reemission_LTE<SM>(m, r, idx, p_icell, rand2, lte_prefetch(m, idx));
  else if ((double)sel <= __ldg(g.P_LTE_p_nLTE + cl)) o.lambda = im_reemission_NLTE<BANK>(idx, lambda0, rand1, rand2);
  else o.lambda = im_reemission_qRE<BANK>(idx, lambda0, rand1, rand2);
  return o;
}

// ---- output.f90:294-595 capteur, SED branch ------------------------------------
template <int BANK>
__device__ __noinline__ int capteur(int lambda, double u1, double v1, double w1,
                                       const double* Sin, bool flag_star, bool flag_scatt) {
  const DevModel& m = c_m; const DevRun& r = c_r;
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
  if (r.lmono0) return capt;          // lmono0 .and. .not.loutput_mc: no MC map is kept (output.f90:360)
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

// ---- output.f90:294-595 capteur, complete: packet origin (lorigine), lonly_capt_interet, and the Monte
// Carlo photon maps of the image step (rotation to the observer's frame, disk position angle, pixel,
// left-right half-photon symmetry).  x,y,z is the START of the last flight (physical_length leaves its inout
// position untouched when the packet exits, optical_depth.f90:86-90), idx0 the 0-based cell it started in.
#define POS0(k, slot) (c_m.pos0[((size_t)blockIdx.x * 4 + (k)) * NP + (slot)])
template <int BANK>
__device__ __noinline__ int capteur_full(int lambda, double x1, double y1, double z1, double u1, double v1, double w1,
                                         const double* Sin, bool flag_star, bool flag_scatt, int idx0) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  double s0 = Sin[0], s1 = Sin[1], s2 = Sin[2], s3 = Sin[3];
  if (w1 < 0.0) {
    if (r.l_sym_centrale) { x1 = -x1; y1 = -y1; z1 = -z1; u1 = -u1; v1 = -v1; w1 = -w1; s2 = -s2; }
    else return 0;
  }
  int capt = (int)((-1.0 * w1 + 1.0) * r.N_thet) + 1;
  if (capt == r.N_thet + 1) capt = r.N_thet;
  if (r.lorigine && capt == r.capt_interet) {
    if (flag_star) atomicAdd(m.star_origin + (lambda - 1), s0);
    else if (idx0 >= 0) atomicAdd(m.disk_origin + (lambda - 1) + (size_t)m.n_lambda * idx0, s0);
  }
  if (r.lmono0 && !r.mc_maps) return capt;
  if (r.lonly_capt_interet && ((capt > r.capt_sup) || (capt < r.capt_inf))) return capt;
  int c_phi = 1;
  if (r.l_sym_axiale) {
    if (v1 < 0.0) { v1 = -v1; y1 = -y1; s2 = -s2; }
    if (w1 != 1.0) c_phi = (int)(atan2(v1, u1) / MCB_PI * r.N_phi) + 1;
  } else {
    if (w1 != 1.0) c_phi = (int)(fmodulo(atan2(u1, v1) + MCB_PI / 2, 2 * MCB_PI) / (2 * MCB_PI) * r.N_phi) + 1;
  }
  if (c_phi == r.N_phi + 1) c_phi = r.N_phi; else if (c_phi == 0) c_phi = 1;
  if (r.lmono0) {
    const int maxigrid = max(r.npix_x, r.npix_y);
    int deltapix_x = 1, deltapix_y = 1;
    if (r.npix_x > r.npix_y) deltapix_y = 1 - (r.npix_x / 2) + (r.npix_y / 2);
    else if (r.npix_x < r.npix_y) deltapix_x = 1 - (r.npix_y / 2) + (r.npix_x / 2);
    const double size_pix = maxigrid / r.map_size;
    double xprim, yprim, zprim;
    rotation(x1, y1, z1, u1, v1, w1, xprim, yprim, zprim);
    double ytmp = yprim; const double ztmp = zprim;
    yprim = ytmp * r.cos_disk + ztmp * r.sin_disk;
    zprim = ztmp * r.cos_disk - ytmp * r.sin_disk;
    const int imap1 = (int)((yprim * r.zoom + 0.5 * r.map_size) * size_pix) + deltapix_x;
    if (imap1 <= 0 || imap1 > r.npix_x) return capt;
    const int jmap1 = (int)((zprim * r.zoom + 0.5 * r.map_size) * size_pix) + deltapix_y;
    if (jmap1 <= 0 || jmap1 > r.npix_y) return capt;
    const size_t plane = (size_t)r.npix_x * r.npix_y * r.N_thet * r.N_phi;
    const int i_contrib = r.n_stokes + (flag_star ? (flag_scatt ? 1 : 0) : (flag_scatt ? 3 : 2));
    auto add = [&](int im, int jm, double f, double sign_u) {
      const size_t q = (size_t)(im - 1) + (size_t)r.npix_x * ((size_t)(jm - 1) + (size_t)r.npix_y * ((size_t)(capt - 1) + (size_t)r.N_thet * (c_phi - 1)));
      atomicAdd(m.smap + q, f * s0);
      if (r.lsepar_pola) { atomicAdd(m.smap + plane + q, f * s1); atomicAdd(m.smap + 2 * plane + q, sign_u * f * s2); atomicAdd(m.smap + 3 * plane + q, f * s3); }
      if (r.lsepar_contrib) atomicAdd(m.smap + (size_t)i_contrib * plane + q, f * s0);
    };
    if (r.l_sym_ima) {
      add(imap1, jmap1, 0.5, 1.0);
      ytmp = -ytmp;
      yprim = ytmp * r.cos_disk - ztmp * r.sin_disk;
      zprim = ztmp * r.cos_disk + ytmp * r.sin_disk;
      const int imap2 = (int)((yprim * r.zoom + 0.5 * r.map_size) * size_pix) + deltapix_x;
      if (imap2 <= 0 || imap2 > r.npix_x) return capt;
      const int jmap2 = (int)((zprim * r.zoom + 0.5 * r.map_size) * size_pix) + deltapix_y;
      if (jmap2 <= 0 || jmap2 > r.npix_y) return capt;
      if (imap1 == imap2 && jmap1 == jmap2) add(imap1, jmap1, 0.5, 1.0);
      else add(imap2, jmap2, 0.5, -1.0);
    } else add(imap1, jmap1, 1.0, 1.0);
    return capt;
  }
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

template <int BANK>
__device__ __noinline__ void angles_scatt_rt1(double u, double v, double w, Rt1Scratch& sc) {
  const DevRun& r = c_r;
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
template <int BANK>
__device__ __noinline__ void deposit_rt1(int idx, int p_icell, int p_lambda, double l,
                                            const double* S, bool flag_star, double xm, double ym, double zm, const Rt1Scratch& sc) {
  const DevModel& m = c_m; const DevRun& r = c_r;
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

// ---- radiation_field.f90:91-130: rt2 specific intensity I_spec (2D only), fp32 reductions ----
template <int BANK>
__device__ __noinline__ void deposit_rt2(int idx, double l, const double* S, bool flag_star, bool flag_direct_star,
                                         double xm, double ym, double zm, double u, double v, double w) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  if (flag_direct_star) { atomicAdd(m.I_spec_star + idx, (float)(l * S[0])); return; }
  const double phi_pos = atan2(xm, ym);
  const double phi_vol = atan2(-u, -v) + MCB_TWO_PI;
  int phi_I = (int)floor(fmodulo(phi_vol - phi_pos, MCB_TWO_PI) / MCB_TWO_PI * r.n_phi_I) + 1;
  if (phi_I > r.n_phi_I) phi_I = 1;
  int theta_I = (int)floor(0.5 * ((zm > 0.0 ? w : -w) + 1.0) * r.n_theta_I) + 1;
  if (theta_I > r.n_theta_I) theta_I = r.n_theta_I;
  float* base = m.I_spec + (size_t)r.n_type_flux * ((size_t)(theta_I - 1) + (size_t)r.n_theta_I * ((size_t)(phi_I - 1) + (size_t)r.n_phi_I * (size_t)idx));
  for (int is = 0; is < r.n_stokes; ++is) atomicAdd(base + is, (float)(l * S[is]));
  if (r.lsepar_contrib) atomicAdd(base + r.n_stokes + (flag_star ? 1 : 3), (float)(l * S[0]));
}

// =============================================================================
// Packet pool in shared memory
// =============================================================================
// 768 threads (24 warps, 80 registers) per block: measured steady-state packets/s on the ref4.1-like thermal step
// 512: 7.98e7, 576: 8.56e7, 640: 9.19e7, 704: 9.22e7, 768: 9.74e7, 832: 8.99e7, 896: 9.34e7, 1024: 9.10e7
#ifndef MCB_BLOCK_T
#define MCB_BLOCK_T 768
#endif
// cell crossings per FLY visit (a visit also ends when fewer than half of its packets still fly): steady-state
// packets/s at 768 threads  5: 9.42e7, 8: 9.73e7, 12: 9.99e7, 16: 1.006e8
#ifndef MCB_FLY_STEPS_T
#define MCB_FLY_STEPS_T 16
#endif
constexpr int MC_BLOCK = MCB_BLOCK_T;       // threads per block (one block per SM)
constexpr int NP = 1024;            // packets in flight per block
constexpr int FLY_STEPS = MCB_FLY_STEPS_T;        // max cell crossings per FLY visit
constexpr unsigned DRAIN_LIVE = 96; // live packets per block below which the pool is considered to be draining out
// Straggler hand-over (DevRun.park_enable): once the packet counter is dry and at most DevRun.park_live (<= PARK_LIVE)
// packets of a block are still in flight, the block writes them to a global buffer and exits; a second, small launch
// of the same kernel (adopt = 1) takes them from that buffer instead of emitting new packets and finishes them on a
// few SMs, while the next call's main launch (another handle, another stream) already owns the rest of the GPU.
// Why: when the counter runs dry, ~15 % of the slots hold packets of the longest-lived 0.1 % (an in-flight packet is
// a length-biased sample); their remaining events are ~30 ms worth of throughput but ~1 s of dependent latency.
constexpr unsigned PARK_LIVE = 256;
constexpr int PARK_REC = 20;        // doubles per parked packet: 11 F fields | 10 u32 (9 U fields, queue) | Q,U,V | pad

enum { F_PX = 0, F_PY, F_PZ, F_OX, F_OY, F_OZ, F_U, F_V, F_W, F_S0, F_EXTR };
// Stokes Q,U,V of packet `slot` of this block: quv[(blockIdx.x*3 + k)*NP + slot]
#define QUV(k, slot) (c_m.quv[((size_t)blockIdx.x * 3 + (k)) * NP + (slot)])
enum { U_C0A = 0, U_C0B, U_COA, U_COB, U_PKLO, U_PKHI, U_EV, U_MISC, U_RALB, NU32 = 9 };
// Stokes Q,U,V live in GLOBAL memory (DevModel.quv, one L2-resident slab per block): they are only touched by
// scatterings, re-emissions and the detector, and keeping them out of shared memory leaves more of the
// 227 KB for L1 (kappa_factor / volume / tally lines).
__host__ __device__ constexpr int pool_nf64(bool) { return 11; }

// bytes of shared memory after the staged tables
__host__ __device__ constexpr size_t pool_bytes(bool pola) {
  return (size_t)pool_nf64(pola) * NP * 8 + (size_t)NU32 * NP * 4 + (size_t)NQ * NP * 2 + 64 * 4;
}

struct Pool {
  double* f;            // [NF64][NP]
  uint32_t* u;          // [NU32][NP]
  unsigned short* q;    // [NQ][NP] ring buffers of slot ids (0xFFFF = entry not written yet)
  unsigned* ctl;        // heads, tails, live packets, ... (CTL_* above)
  __device__ __forceinline__ double& F(int field, int slot) const { return f[field * NP + slot]; }
  __device__ __forceinline__ uint32_t& U(int field, int slot) const { return u[field * NP + slot]; }
  __device__ __forceinline__ volatile unsigned short* Q(int queue) const { return q + queue * NP; }
  __device__ __forceinline__ volatile unsigned& HEAD(int queue) const { return ctl[queue]; }
  __device__ __forceinline__ volatile unsigned& TAIL(int queue) const { return ctl[NQ + queue]; }
  __device__ __forceinline__ volatile unsigned& LIVE() const { return ctl[CTL_LIVE]; }
  __device__ __forceinline__ volatile unsigned& BUSY() const { return ctl[CTL_BUSY]; }
  __device__ __forceinline__ volatile unsigned& PARK() const { return ctl[CTL_PARK]; }   // this block is handing its packets over
  __device__ __forceinline__ volatile unsigned& DRYF() const { return ctl[CTL_DRY]; }   // the global packet counter ran dry
  __device__ __forceinline__ volatile unsigned& SENT() const { return ctl[CTL_SENT]; }   // latest value of the global packet counter seen by this block (saturated)
};

// misc word: lambda (13 bits) | star 1 | scatt 1 | ISM 1 | i_star_hit 4 | n_iteractions_in_cell 8 (saturating) | 4 spare.
// (the chunk of a packet is recovered from its id: packet = (nnfot1 - 1) << 40 | index, see misc_chunk)
__device__ __forceinline__ uint32_t pack_misc(int lambda, bool star, bool scatt, bool ism, int istar, int n_in_cell) {
  return (uint32_t)lambda | ((uint32_t)star << 13) | ((uint32_t)scatt << 14) | ((uint32_t)ism << 15) | ((uint32_t)istar << 16) | ((uint32_t)n_in_cell << 20);
}
constexpr uint32_t MISC_SCATT = 1u << 14;
__device__ __forceinline__ int misc_lambda(uint32_t m) { return m & 8191; }
__device__ __forceinline__ bool misc_star(uint32_t m) { return (m >> 13) & 1; }
__device__ __forceinline__ bool misc_scatt(uint32_t m) { return (m >> 14) & 1; }
__device__ __forceinline__ bool misc_ism(uint32_t m) { return (m >> 15) & 1; }
__device__ __forceinline__ int misc_istar(uint32_t m) { return (m >> 16) & 15; }
__device__ __forceinline__ int misc_n_in_cell(uint32_t m) { return (m >> 20) & 255; }
__device__ __forceinline__ uint32_t misc_set_istar(uint32_t m, int istar) { return (m & ~(15u << 16)) | ((uint32_t)istar << 16); }
__device__ __forceinline__ uint32_t misc_set_n_in_cell(uint32_t m, int n) { return (m & ~(255u << 20)) | ((uint32_t)(n > 255 ? 255 : n) << 20); }
// local chunk index of a packet from the high word of its id (nnfot1 - 1 = pk_hi >> 8)
template <int BANK> __device__ __forceinline__ int chunk_of(uint32_t pk_hi) {
  const DevRun& r = c_r;
  const int nnfot1 = (int)(pk_hi >> 8) + 1;
  const int first_local = r.nnfot1_start + ((r.rank - ((r.nnfot1_start - 1) % r.n_ranks) + r.n_ranks) % r.n_ranks);
  return (nnfot1 - first_local) / r.n_ranks;
}

// cells <-> two 32-bit words
__device__ __forceinline__ void pack_cell(Cell c, uint32_t& a, uint32_t& b) { a = (uint32_t)c.ri; b = ((uint32_t)c.zj & 0xFFFFu) | ((uint32_t)c.k << 16); }
__device__ __forceinline__ void unpack_cell(uint32_t a, uint32_t b, Cell& c) { c.ri = (int)a; c.zj = (int)(short)(b & 0xFFFFu); c.k = (int)(b >> 16); }
__device__ __forceinline__ void pack_cell(int c, uint32_t& a, uint32_t& b) { a = (uint32_t)c; b = 0; }
__device__ __forceinline__ void unpack_cell(uint32_t a, uint32_t, int& c) { c = (int)a; }

// the pool of this block: right after the staged tables in dynamic shared memory
template <bool SM, int BANK> __device__ __forceinline__ Pool make_pool() {
  Pool P;
  unsigned char* base = mcb_smem_raw + (SM ? (size_t)c_m.sm.total_words * 8 : 0);
  const size_t nf = (size_t)pool_nf64(c_r.lsepar_pola != 0);
  P.f = reinterpret_cast<double*>(base);
  P.u = reinterpret_cast<uint32_t*>(base + nf * NP * 8);
  P.q = reinterpret_cast<unsigned short*>(base + nf * NP * 8 + (size_t)NU32 * NP * 4);
  P.ctl = reinterpret_cast<unsigned*>(base + nf * NP * 8 + (size_t)NU32 * NP * 4 + (size_t)NQ * NP * 2);
  return P;
}

struct Stats { unsigned int pk, steps, inter, sca, abs_, kill, esc, bounce, mrw_w, mrw_s; };
// scheduling diagnostics (per warp, lane 0): chunk visits and valid lanes per phase
struct SchedStats { unsigned int visits[NQ], lanes[NQ]; };      // (the first four are reported: EMIT, ABSORB, SCATTER, FLY)

// Regrouping step: push the packets of this warp to the ring queue of their next phase (one shared-
// memory atomic per destination queue).  Packets that leave the pool (no more work) decrement LIVE.
// The fence orders the packet-state stores before the queue entry becomes visible to a consumer.
__device__ __forceinline__ void push_next(const Pool& P, int slot, int nextq, bool valid, unsigned lane) {
  __threadfence_block();
#pragma unroll
  for (int qi = 0; qi < NQ; ++qi) {
    const unsigned mask = __ballot_sync(0xffffffffu, nextq == qi);
    if (mask == 0) continue;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atomicAdd((unsigned*)&P.ctl[NQ + qi], (unsigned)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (nextq == qi) P.Q(qi)[(base + __popc(mask & ((1u << lane) - 1u))) & (NP - 1)] = (unsigned short)slot;
  }
  const unsigned gone = __ballot_sync(0xffffffffu, valid && nextq == Q_NONE);
  if (gone && lane == 0) atomicSub((unsigned*)&P.ctl[CTL_LIVE], (unsigned)__popc(gone));
}

// ---- begin flight `ev`: tau and the interaction-type draw from block 2*ev (dust_transfer.f90:1208-1215,1280),
// physical_length preamble (optical_depth.f90:53-68).  Position / direction / cell are already in the pool. ----
__device__ __forceinline__ float tau_of_rand(float rand) {
  if (rand == 1.0f) return 1.0e30f;
  if (rand > 1.0e-6f) return -logf(1.0f - rand);       // `real` arithmetic in the reference (dust_transfer.f90:1212)
  return rand;
}
template <int BANK>
__device__ __forceinline__ void start_flight(const Pool& P, int slot, double x, double y, double z, double u, double v, double w,
                                             const uint4 b /* Philox block 2*ev of this packet */, uint32_t& misc) {
  const DevModel& m = c_m;
  P.F(F_EXTR, slot) = (double)tau_of_rand(u01(b.x));
  P.U(U_RALB, slot) = __float_as_uint(u01(b.y));
  P.F(F_OX, slot) = x; P.F(F_OY, slot) = y; P.F(F_OZ, slot) = z;
  P.U(U_COA, slot) = 0xFFFFFFF9u; P.U(U_COB, slot) = 0;          // null previous cell
  const int istar = intersect_stars(m, x, y, z, u, v, w);
  misc = misc_set_istar(misc, istar);
}

// =============================================================================
// Modified random walk (MRW.f90, call site dust_transfer.f90:1222-1239; Min et al. 2009, Robitaille 2010).
// The walk runs in the packet-per-warp kernel only (WARP = true; the packet-per-lane kernel skips it: a skipped walk is
// made of ordinary flights, and inside a 32-packet warp it stalled 31 packets for ~10 us).  There all
// 32 lanes run this for the SAME packet: mutable global memory is read by lane 0 and broadcast (so that the lanes stay
// bit-identical) and only lane 0 deposits.
// `ev` is the number of the flight about to start; every step and the closing re-emission take one event number each,
// Philox block (2 ev + 1) | 0x40000000.
// =============================================================================
template <bool WARP> __device__ __forceinline__ double bcast0(double v) { return WARP ? __shfl_sync(0xffffffffu, v, 0) : v; }
template <bool WARP> __device__ __forceinline__ int bcast0(int v) { return WARP ? __shfl_sync(0xffffffffu, v, 0) : v; }
template <bool WARP> __device__ __forceinline__ LtePre lte_prefetch_w(const DevModel& m, int idx) {
  LtePre p = lte_prefetch(m, idx);
  p.xkj = bcast0<WARP>(p.xkj); p.Ti = bcast0<WARP>(p.Ti);
  return p;
}
// MRW.f90:58-70 sample_zeta = interp(y_MRW, zeta, zeta_random) (utils.f90:190-247): first j in 2..n-1 with zeta(j) > xp, else n
__device__ __forceinline__ double sample_zeta(const DevModel& m, double zr) {
  int a = 1, b = N_ZETA - 1;
  while (a < b) { const int mid = (a + b) >> 1; if (__ldg(m.zeta + mid) > zr) b = mid; else a = mid + 1; }
  const double x0 = __ldg(m.zeta + a - 1), x1 = __ldg(m.zeta + a);
  const double y0 = (double)(a - 1) / (double)(N_ZETA - 1), y1 = (double)a / (double)(N_ZETA - 1);
  const double frac = (zr - x0) / (x1 - x0);
  return y0 * (1. - frac) + y1 * frac;
}
struct MrwOut { double x, y, z, u, v, w; int lambda; uint32_t ev; unsigned steps; };
// Cheap pre-test of a walk: distance to the closest wall against the mean free path this cell had the last time a walk
// was evaluated in it (0 = never: try).  Skipping an attempt is harmless -- the packet just makes ordinary flights, which
// is what the walk stands for -- so a stale value only costs or saves time.  Most attempts fail (a packet that keeps
// interacting in one cell usually sits near a wall of a moderately thick cell), and this keeps them at a few instructions.
template <class G, int BANK>
__device__ __forceinline__ bool mrw_worth_trying(typename G::CellT cell, int idx, double x, double y, double z) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  const float2 c = __ldcg(reinterpret_cast<const float2*>(m.mrw_lR) + idx);      // (mean free path, temperature index it was evaluated at)
  if (c.x == 0.0f || (int)c.y != __ldcg(m.xT_ech + idx)) return true;           // never evaluated, or the cell has warmed up since
  return G::closest_wall(m, cell, x, y, z) > 0.8 * r.gamma_MRW * (double)c.x;
}
template <class G, bool SM, int BANK, bool WARP>
__device__ __noinline__ MrwOut mrw_walk(typename G::CellT cell, int idx, int p_icell, double x, double y, double z, double S0,
                                        uint32_t pk_lo, uint32_t pk_hi, uint32_t ev) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  const bool leader = !WARP || (threadIdx.x & 31) == 0;
  MrwOut o; o.x = x; o.y = y; o.z = z; o.u = 0; o.v = 0; o.w = 1; o.lambda = 0; o.ev = ev; o.steps = 0;
  const double kf = __ldg(m.kappa_factor + idx);
  if (!(kf > 0.0)) return o;
  double d = G::closest_wall(m, cell, x, y, z);
  int Ti; double frac_T2;
  temp_lte<SM>(m, r, idx, p_icell, lte_prefetch_w<WARP>(m, idx), Ti, frac_T2);
  const double frac_T1 = 1.0 - frac_T2;
  const size_t q1 = (size_t)(Ti - 2) + (size_t)m.n_T * (p_icell - 1), q2 = q1 + 1;
  const double A = frac_T1 * __ldg(m.mrw_A + q1) + frac_T2 * __ldg(m.mrw_A + q2);
  const double B = frac_T1 * __ldg(m.mrw_B + q1) + frac_T2 * __ldg(m.mrw_B + q2);
  const double Cc = frac_T1 * __ldg(m.mrw_C + q1) + frac_T2 * __ldg(m.mrw_C + q2);
  if (!(A > 0.0) || !(B > 0.0)) return o;
  const double l_R = B / (A * kf);                  // Rosseland-type mean free path 1 / (rho chi_R)
  if (leader) reinterpret_cast<float2*>(m.mrw_lR)[idx] = make_float2((float)l_R, (float)Ti);      // remembered per cell: the cheap pre-test of the next attempts (mrw_worth_trying)
  while (d > r.gamma_MRW * l_R && o.steps < 100000u) {
    const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), (2u * o.ev + 1u) | 0x40000000u, pk_lo, pk_hi, r.call_index);
    double u, v, w;
    random_isotropic_direction(u01(b.x), u01(b.y), u, v, w);      // MRW.f90:84-88: random point on the sphere of radius d
    o.x = o.x + u * d; o.y = o.y + v * d; o.z = o.z + w * d;
    double zr = (double)u01(b.z);
    if (zr <= 0.0) zr = 1.0 / 33554432.0;
    const double ym = sample_zeta(m, zr);                          // MRW.f90:92
    const double ct = -mcb_log(ym) * (d / MCB_PI) * (d / MCB_PI) * 3.0 / l_R;     // Min et al. 2009 eq. 8 with D = l_R / 3
    if (leader) atomicAdd(m.tally + m.lay.xKJ + idx, S0 * ct * Cc / A);            // energy left along the walk
    ++o.steps; ++o.ev;
    d = G::closest_wall(m, cell, o.x, o.y, o.z);
  }
  if (o.steps == 0u) return o;
  // end of the walk: thermal re-emission at the cell's running temperature, isotropic, unpolarised
  const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), (2u * o.ev + 1u) | 0x40000000u, pk_lo, pk_hi, r.call_index);
  if (WARP) __syncwarp();
  o.lambda = im_reemission_LTE<SM>(m, r, idx, p_icell, u01(b.y), lte_prefetch_w<WARP>(m, idx));
  random_isotropic_direction(u01(b.z), u01(b.w), o.u, o.v, o.w);
  ++o.ev;
  return o;
}

// =============================================================================
// emit_packet (dust_transfer.f90:1047-1151) of packet (pk_lo, pk_hi): wavelength (thermal / polychromatic calls), source,
// position, direction, first cell.  Pure function of the packet id (Philox blocks 0 and 1), shared by the packet-per-lane
// kernel (phase_emit) and the packet-per-warp kernel (warp_engine.cuh).
// =============================================================================
template <class CellT> struct Emitted { double x, y, z, u, v, w, S0; CellT cell; int lambda; bool flag_star, flag_ISM, lintersect; };

template <class G, bool SM, int BANK, int VAR>
__device__ __forceinline__ Emitted<typename G::CellT> emit_packet_core(uint32_t pk_lo, uint32_t pk_hi) {
  constexpr bool GR = VAR == VAR_EXTRAS; constexpr bool TH = VAR == VAR_THERMAL; (void)GR; (void)TH;
  const DevModel& m = c_m; const DevRun& r = c_r;
  using CellT = typename G::CellT;
  Emitted<CellT> e;
  const uint4 b0 = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 0u, pk_lo, pk_hi, r.call_index);
  const uint4 b1 = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 1u, pk_lo, pk_hi, r.call_index);
  int di = 0;
  auto nextf = [&]() -> float {
    const int i = di++;
    const uint32_t w = (i == 0) ? b0.x : (i == 1) ? b0.y : (i == 2) ? b0.z : (i == 3) ? b0.w : (i == 4) ? b1.x : (i == 5) ? b1.y : (i == 6) ? b1.z : b1.w;
    return u01(w);
  };
  // n_phot_envoyes(lambda) is incremented with the PREVIOUS packet's lambda in thermal mode
  // (dust_transfer.f90:531 precedes :537); on the device packets are unordered, so the count is
  // attributed to the packet's own emission wavelength (the sum over lambda is identical).
  int lambda = r.lambda_in;
  if (TH || !r.lmono) lambda = select_wl_em<SM>(m, nextf());
  e.lambda = lambda;
  double x, y, z, u, v, w, S0;
  CellT cell; null_cell(cell);
  bool lintersect = true;
  float rand = nextf();
  const bool ism_only = !TH && r.lism;      // ISM side loop: emit_packet_ISM for every packet (dust_transfer.f90:967)
  if (!ism_only && (double)rand <= t_frac_star<SM>(m, lambda)) {
    e.flag_star = true; e.flag_ISM = false;
    const int i_star = select_star(m, lambda, nextf());
    const float rand1 = nextf(), rand2 = nextf(), rand3 = nextf(), rand4 = nextf();
    // emit_packet_uniform_sphere (stars.f90:108-169)
    double zz = 2.0 * rand1 - 1.0;
    double srw02 = sqrt(1.0 - zz * zz), sa, ca;
    mcb_sincospi(2.0 * rand2 - 1.0, &sa, &ca);            // argmt = pi*(2 rand2 - 1)
    double xx = srw02 * ca, yy = srw02 * sa;
    double cospsi = (double)sqrtf(rand3), sp, cp;
    mcb_sincospi(2.0 * (double)rand4, &sp, &cp);          // phi = 2 pi rand4
    cdapres(cospsi, sp, cp, xx, yy, zz, u, v, w);
    const double r_star = m.star[i_star - 1][3] * (1.0 + 1e-6);
    x = xx * r_star + m.star[i_star - 1][0]; y = yy * r_star + m.star[i_star - 1][1]; z = zz * r_star + m.star[i_star - 1][2];
    if (G::is_vor) cell_of_id(m, m.star_icell[i_star - 1], cell);
    else cell = G::index(m, x, y, z);
    if (m.star_out[i_star - 1]) lintersect = G::move_to_grid(m, x, y, z, u, v, w, cell);
    S0 = m.E_paquet;
    if (GR && r.lspot) {      // hot spot on star 1 (dust_transfer.f90:1094-1119), tested on the position emit_packet_uniform_sphere returns
      if ((double)r.x_spot * x + (double)r.y_spot * y + (double)r.z_spot * z > (double)r.cos_thet_spot * m.star[0][3]) {
        const float hc_lk = (float)(6.626070040e-34 * 299792458.0 / (__ldg(m.tab_lambda + lambda - 1) * 1e-6 * 1.38064852e-23));
        const float correct_spot = (float)((exp((double)hc_lk / r.star1_T) - 1) / (double)(expf(hc_lk / r.T_spot) - 1));
        S0 = S0 * correct_spot;
      }
    }
  } else if (!ism_only && (double)rand <= t_frac_disk<SM>(m, lambda)) {
    e.flag_star = false; e.flag_ISM = false;
    const int ic = select_cellule(m, lambda, nextf());
    cell_of_id(m, ic, cell);
    const float rand1 = nextf(), rand2 = nextf(), rand3 = nextf();
    G::pos_em_cell(m, cell, rand1, rand2, rand3, x, y, z);
    const float rw = nextf(), rp = nextf();
    random_isotropic_direction(rw, rp, u, v, w);
    S0 = m.E_paquet;
    if (GR && r.lweight_emission) S0 = S0 * __ldg(m.correct_E + ic - 1);      // dust_transfer.f90:1140-1142
  } else {
    e.flag_star = false; e.flag_ISM = true;
    // emit_packet_ISM (stars.f90:728-787)
    S0 = 1.0;
    const float rand1 = nextf(), rand2 = nextf();
    double zz = 2.0 * rand1 - 1.0;
    double srw02 = sqrt(1.0 - zz * zz), sa, ca;
    mcb_sincospi(2.0 * rand2 - 1.0, &sa, &ca);
    double xx = srw02 * ca, yy = srw02 * sa;
    const float rand3 = nextf(), rand4 = nextf();
    double cospsi = (double)(-sqrtf(rand3)), sp, cp;
    mcb_sincospi(2.0 * (double)rand4, &sp, &cp);
    cdapres(cospsi, sp, cp, xx, yy, zz, u, v, w);
    x = m.cISM[0] + xx * m.R_ISM; y = m.cISM[1] + yy * m.R_ISM; z = m.cISM[2] + zz * m.R_ISM;
    lintersect = G::move_to_grid(m, x, y, z, u, v, w, cell);
  }
  e.x = x; e.y = y; e.z = z; e.u = u; e.v = v; e.w = w; e.S0 = S0; e.cell = cell; e.lintersect = lintersect;
  return e;
}

// =============================================================================
// EMIT: claim a packet id, emit_packet (dust_transfer.f90:1047-1151), start the first flight
// =============================================================================
template <class G, bool SM, int BANK, int VAR>
__device__ __noinline__ int phase_emit(int slot, bool valid, Stats& st) {
  constexpr bool GR = VAR == VAR_EXTRAS; constexpr bool TH = VAR == VAR_THERMAL; (void)GR; (void)TH;
  const DevModel& m = c_m; const DevRun& r = c_r;
  const bool POLA = r.lsepar_pola != 0;
  const Pool P = make_pool<SM, BANK>();
  using CellT = typename G::CellT;
  const unsigned lane = threadIdx.x & 31;
  int nextq = Q_NONE;
  // ---- claim.  Thermal / fixed-count mode: one warp-aggregated atomic on the global counter.  SED mode
  // (dust_transfer.f90:507-510,529,551): a chunk keeps sending until n_photons2 packets were RECEIVED in
  // detector bin capt_sup (or n_phot_lim were sent).
  int nnfot1 = 0, my_chunk = 0;
  unsigned long long idx_in_chunk = 0;
  bool got = false;
  const int first_local = r.nnfot1_start + ((r.rank - ((r.nnfot1_start - 1) % r.n_ranks) + r.n_ranks) % r.n_ranks);
  if (TH || r.count_sent) {
    const unsigned need = __ballot_sync(0xffffffffu, valid);
    if (need) {
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(m.work, (unsigned long long)__popc(need));
      base = __shfl_sync(0xffffffffu, base, leader);
      const unsigned long long g = base + __popc(need & ((1u << lane) - 1u));
      if ((int)lane == leader) P.SENT() = (unsigned)(base > 0xffffffffull ? 0xffffffffull : base);
      if (valid && g >= r.n_packets_total) { atomicMin(m.work + 1, (unsigned long long)globaltimer_ns()); P.DRYF() = 1u; }     // start of the drain-out
      if (valid && g < r.n_packets_total) {
        const unsigned long long lc = g / r.n_per_chunk;
        idx_in_chunk = g % r.n_per_chunk;
        nnfot1 = first_local + (int)lc * r.n_ranks;
        my_chunk = (int)lc;
        got = true;
      }
    }
  } else if (valid) {
    const int start = (int)(((unsigned long long)blockIdx.x * NP + slot + st.pk) % (unsigned long long)r.n_local_chunks);
    for (int tries = 0; tries < r.n_local_chunks && !got; ++tries) {
      int lc = start + tries; if (lc >= r.n_local_chunks) lc -= r.n_local_chunks;
      if (__ldcg(m.work + 3 + 2 * lc) >= (unsigned long long)r.n_photons2) continue;
      const unsigned long long sidx = atomicAdd(m.work + 2 + 2 * lc, 1ull);
      if (sidx >= r.sent_lim) { atomicAdd(m.work + 2 + 2 * lc, ~0ull); continue; }     // undo (adds -1)
      idx_in_chunk = sidx; nnfot1 = first_local + lc * r.n_ranks; my_chunk = lc; got = true;
    }
  }
  if (got) {
    const unsigned long long packet = ((unsigned long long)(nnfot1 - 1) << 40) + idx_in_chunk;
    const uint32_t pk_lo = (uint32_t)packet, pk_hi = (uint32_t)(packet >> 32);
    ++st.pk;
    const Emitted<CellT> e = emit_packet_core<G, SM, BANK, VAR>(pk_lo, pk_hi);
    const int lambda = e.lambda;
    atomicAdd(m.tally + m.lay.n_env + (lambda - 1), 1.0);
    if (!TH && r.lism && e.lintersect) atomicAdd(m.work + 3 + 2 * my_chunk, 1ull);      // nnfot2 counts the packets that enter the model (:973-976)
    if (e.lintersect) {
      P.F(F_PX, slot) = e.x; P.F(F_PY, slot) = e.y; P.F(F_PZ, slot) = e.z;
      P.F(F_U, slot) = e.u; P.F(F_V, slot) = e.v; P.F(F_W, slot) = e.w;
      P.F(F_S0, slot) = e.S0;
      if (POLA) { QUV(0, slot) = 0.0; QUV(1, slot) = 0.0; QUV(2, slot) = 0.0; }
      uint32_t ca_, cb_; pack_cell(e.cell, ca_, cb_);
      P.U(U_C0A, slot) = ca_; P.U(U_C0B, slot) = cb_;
      P.U(U_PKLO, slot) = pk_lo; P.U(U_PKHI, slot) = pk_hi; P.U(U_EV, slot) = 1u;
      uint32_t misc = pack_misc(lambda, e.flag_star, false, e.flag_ISM, 0, 0);
      start_flight<BANK>(P, slot, e.x, e.y, e.z, e.u, e.v, e.w, philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u, pk_lo, pk_hi, r.call_index), misc);
      if (GR && r.capt_full) { POS0(0, slot) = e.x; POS0(1, slot) = e.y; POS0(2, slot) = e.z; POS0(3, slot) = (double)tally_index(m, e.cell); }
      P.U(U_MISC, slot) = misc;
      nextq = Q_FLY;
    } else {      // the packet never enters the model: straight to the detector (dust_transfer.f90:545-552)
      if (!e.flag_ISM) {
        const double S[4] = {e.S0, 0.0, 0.0, 0.0};
        const int capt = (GR && r.capt_full) ? capteur_full<BANK>(lambda, e.x, e.y, e.z, e.u, e.v, e.w, S, e.flag_star, false, tally_index(m, e.cell))
                                     : capteur<BANK>(lambda, e.u, e.v, e.w, S, e.flag_star, false);
        if (!TH && !r.count_sent && capt == r.capt_sup) atomicAdd(m.work + 3 + 2 * my_chunk, 1ull);
        ++st.esc;
      }
      nextq = Q_EMIT;
    }
  }
  return nextq;
}

// =============================================================================
// FLY: up to FLY_STEPS iterations of the physical_length loop (optical_depth.f90:77-178)
// =============================================================================
template <class G, bool SM, int BANK, int VAR>
__device__ __noinline__ int phase_fly(int slot, bool valid, Stats& st) {
  constexpr bool GR = VAR == VAR_EXTRAS; constexpr bool TH = VAR == VAR_THERMAL; (void)GR; (void)TH;
  const DevModel& m = c_m; const DevRun& r = c_r;
  const bool POLA = r.lsepar_pola != 0;
  const Pool P = make_pool<SM, BANK>();
  using CellT = typename G::CellT;
  using Hit = typename G::Hit;
  const bool thermal = TH || r.letape_th != 0;
  const bool variable_dust = !SM && m.p_n_cells != 1;      // staged tables imply p_n_cells == 1
  const bool rt1_on = !TH && (!thermal) && r.rt1;
  int nextq = Q_NONE;
  double x0 = 0, y0 = 0, z0 = 0, u = 0, v = 0, w = 1, extr = 0, S0 = 0, xo = 0, yo = 0, zo = 0;
  uint32_t misc = 0;
  int lambda = 1, i_star_hit = 0;
  CellT c0, c_old; null_cell(c0); null_cell(c_old);
  DirInv dinv; dinv.inv_a = 0; dinv.inv_w = 0;
  Rt1Scratch rt1;
  if (valid) {
    x0 = P.F(F_PX, slot); y0 = P.F(F_PY, slot); z0 = P.F(F_PZ, slot);
    u = P.F(F_U, slot); v = P.F(F_V, slot); w = P.F(F_W, slot);
    extr = P.F(F_EXTR, slot);
    S0 = P.F(F_S0, slot);
    misc = P.U(U_MISC, slot);
    lambda = misc_lambda(misc); i_star_hit = misc_istar(misc);
    unpack_cell(P.U(U_C0A, slot), P.U(U_C0B, slot), c0);
    unpack_cell(P.U(U_COA, slot), P.U(U_COB, slot), c_old);
    xo = P.F(F_OX, slot); yo = P.F(F_OY, slot); zo = P.F(F_OZ, slot);
    dinv = dir_invariants(u, v, w);
    if (rt1_on) angles_scatt_rt1<BANK>(u, v, w, rt1);       // recomputed per visit (same values as once per flight)
    nextq = Q_FLY;
  }
  int idx_c = valid ? tally_index(m, c0) : -1;      // tally index of c0 (-1: virtual cell), updated wherever c0 changes
  CellT c_star; null_cell(c_star);
  if (i_star_hit > 0) cell_of_id(m, m.star_icell[i_star_hit - 1], c_star);      // the cell of the star this flight points at
  const int n_in = __popc(__ballot_sync(0xffffffffu, valid));
  bool flying = valid, interact = false;
#pragma unroll 1
  for (int it = 0; it < FLY_STEPS; ++it) {
    // leave the loop once fewer than half of this chunk's packets are still in flight: the rest of the
    // warp would idle; the packets still flying are re-queued and regrouped into full chunks
    const int n_fly = __popc(__ballot_sync(0xffffffffu, flying));
    if (n_fly == 0 || (it > 0 && 2 * n_fly < n_in)) break;
    if (!flying) continue;
    if (idx_c < 0 && G::test_exit(m, c0, x0, y0, z0)) {      // (a real cell is never an exit)
      if (!misc_ism(misc)) {       // the packet leaves the model: detector (capteur, output.f90:294)
        double S[4] = {S0, 0.0, 0.0, 0.0};
        if (POLA) { S[1] = QUV(0, slot); S[2] = QUV(1, slot); S[3] = QUV(2, slot); }
        const int capt = (GR && r.capt_full) ? capteur_full<BANK>(lambda, POS0(0, slot), POS0(1, slot), POS0(2, slot), u, v, w, S, misc_star(misc), misc_scatt(misc), (int)POS0(3, slot))
                                     : capteur<BANK>(lambda, u, v, w, S, misc_star(misc), misc_scatt(misc));
        if (!TH && !r.count_sent && capt == r.capt_sup) atomicAdd(m.work + 3 + 2 * chunk_of<BANK>(P.U(U_PKHI, slot)), 1ull);
        ++st.esc;      // (interstellar packets leave without being detected, dust_transfer.f90:548)
      }
      nextq = Q_EMIT; flying = false;
      continue;
    }
    if (i_star_hit > 0 && same_cell(c0, c_star)) { ++st.kill; nextq = Q_EMIT; flying = false; continue; }     // packet absorbed by the star
    const int idx = idx_c;
    double opacity = 0.0;
    int p_icell = 1;
    if (idx >= 0) {
      p_icell = variable_dust ? idx + 1 : 1;
      const double kf = __ldg(m.kf_dark + idx);      // kappa_factor, sign bit set <=> l_dark_zone (one load for both)
      if (signbit(kf)) {
        // dark-zone bounce (optical_depth.f90:104-112): back to the previous cell's entry point, reversed
        u = -u; v = -v; w = -w;
        c0 = c_old; x0 = xo; y0 = yo; z0 = zo; idx_c = tally_index(m, c0);
        ++st.bounce;
        interact = true; flying = false;
        continue;
      }
      opacity = t_kappa<SM>(m, p_icell, lambda) * kf;
    }
    const Hit h = G::distance(m, dinv, x0, y0, z0, u, v, w, c0, c_old);
    ++st.steps;
    double l_contrib = hit_l_contrib(h), l = h.l;
    const double tau_c = l_contrib * opacity;
    bool lstop = false;
    if (tau_c > extr) {
      lstop = true;
      l_contrib = l_contrib * mc_div(extr, tau_c);
      l = hit_l_void(h) + l_contrib;
    } else extr = extr - tau_c;
    if (idx >= 0) {
      // save_radiation_field (radiation_field.f90:31-135)
      if (thermal) {
        atomicAdd(m.tally + m.lay.xKJ + idx, t_kappa_abs<SM>(m, p_icell, lambda) * l_contrib * S0);
        if (r.lxJ) atomicAdd(m.tally + m.lay.xJ + idx + (size_t)m.n_cells * (lambda - 1), l_contrib * S0);
        if (GR && r.lxN) atomicAdd(m.xN + idx, 1.0);
      } else {
        if (r.lxJ) {
          atomicAdd(m.tally + m.lay.xJ + idx + (size_t)m.n_cells * (lambda - 1), l_contrib * S0);
          if (GR && r.lxN) atomicAdd(m.xN + idx + (size_t)m.n_cells * (lambda - 1), 1.0);
        }
        if (rt1_on) {
          double x1, y1, z1;
          G::exit_point(h, x0, y0, z0, u, v, w, x1, y1, z1);
          double S[4] = {S0, 0.0, 0.0, 0.0};
          if (POLA) { S[1] = QUV(0, slot); S[2] = QUV(1, slot); S[3] = QUV(2, slot); }
          deposit_rt1<BANK>(idx, p_icell, r.p_lambda_in, l_contrib, S, misc_star(misc),
                      0.5 * (x0 + x1), 0.5 * (y0 + y1), 0.5 * (z0 + z1), rt1);
        } else if (!TH && r.rt2) {
          double x1, y1, z1;
          G::exit_point(h, x0, y0, z0, u, v, w, x1, y1, z1);
          double S[4] = {S0, 0.0, 0.0, 0.0};
          if (POLA) { S[1] = QUV(0, slot); S[2] = QUV(1, slot); S[3] = QUV(2, slot); }
          // flag_direct_star == stellar packet that has not interacted yet (dust_transfer.f90:1189-1193,1262)
          deposit_rt2<BANK>(idx, l_contrib, S, misc_star(misc), misc_star(misc) && !misc_scatt(misc),
                            0.5 * (x0 + x1), 0.5 * (y0 + y1), 0.5 * (z0 + z1), u, v, w);
        }
      }
    }
    if (lstop) {
      x0 = x0 + l * u; y0 = y0 + l * v; z0 = z0 + l * w;                               // interaction point
      if (!G::is_vor && m.l3D && m.kind == 1) { c0 = G::index(m, x0, y0, z0); idx_c = tally_index(m, c0); }      // optical_depth.f90:162-165
      interact = true; flying = false;
      continue;
    }
    double x1, y1, z1;
    CellT c1;
    G::advance(m, h, x0, y0, z0, u, v, w, c0, x1, y1, z1, c1);
    xo = x0; yo = y0; zo = z0; c_old = c0;
    x0 = x1; y0 = y1; z0 = z1; c0 = c1; idx_c = tally_index(m, c1);
  }
  if (valid) {
    if (interact) {
      // the flight ended with an interaction at (x0,y0,z0) in cell c0 (dust_transfer.f90:1260-1284)
      ++st.inter;
      if (!TH && r.lmono) nextq = Q_SCAT;       // forced scattering; the dark-zone / energy tests are done in the SCATTER phase
      else {
        const int idx = idx_c;
        const int p_icell = (variable_dust && idx >= 0) ? idx + 1 : 1;
        nextq = (__uint_as_float(P.U(U_RALB, slot)) < t_albedo<SM>(m, p_icell, lambda)) ? Q_SCAT : Q_ABS;
      }
      P.F(F_U, slot) = u; P.F(F_V, slot) = v; P.F(F_W, slot) = w;       // (reversed on a bounce)
    }
    if (nextq != Q_EMIT) {
      P.F(F_PX, slot) = x0; P.F(F_PY, slot) = y0; P.F(F_PZ, slot) = z0;
      uint32_t ca_, cb_; pack_cell(c0, ca_, cb_);
      P.U(U_C0A, slot) = ca_; P.U(U_C0B, slot) = cb_;
      if (nextq == Q_FLY) {
        P.F(F_OX, slot) = xo; P.F(F_OY, slot) = yo; P.F(F_OZ, slot) = zo;
        pack_cell(c_old, ca_, cb_);
        P.U(U_COA, slot) = ca_; P.U(U_COB, slot) = cb_;
        P.F(F_EXTR, slot) = extr;
      }
    }
  }
  return nextq;
}

// =============================================================================
// SCATTER: method 2 (dust_transfer.f90:1318-1351) + start of the next flight
// =============================================================================
template <class G, bool SM, int BANK, int VAR>
__device__ __noinline__ int phase_scatter(int slot, bool valid, Stats& st) {
  constexpr bool GR = VAR == VAR_EXTRAS; constexpr bool TH = VAR == VAR_THERMAL; (void)GR; (void)TH;
  const DevModel& m = c_m; const DevRun& r = c_r;
  const bool POLA = r.lsepar_pola != 0;
  const Pool P = make_pool<SM, BANK>();
  using CellT = typename G::CellT;
  int nextq = Q_NONE;
  if (valid) {
    const bool variable_dust = !SM && m.p_n_cells != 1;      // staged tables imply p_n_cells == 1
    uint32_t misc = P.U(U_MISC, slot);
    const int lambda = misc_lambda(misc);
    CellT cell; unpack_cell(P.U(U_C0A, slot), P.U(U_C0B, slot), cell);
    const int idx = tally_index(m, cell);
    const int p_icell = (variable_dust && idx >= 0) ? idx + 1 : 1;
    bool dead = idx < 0;      // interaction in a virtual cell (only with an inconsistent dark-zone mask): drop the packet
    double S[4] = {P.F(F_S0, slot), 0.0, 0.0, 0.0};
    if (POLA) { S[1] = QUV(0, slot); S[2] = QUV(1, slot); S[3] = QUV(2, slot); }
    if (!TH && r.lmono) {      // forced scattering (dust_transfer.f90:1263-1278)
      if (dead || __ldg(m.dark + idx)) dead = true;
      else {
        const float albedo = t_albedo<SM>(m, p_icell, lambda);
        S[0] *= albedo; S[1] *= albedo; S[2] *= albedo; S[3] *= albedo;
        if (S[0] < (double)(FLT_MIN * 1.0e6f)) dead = true;
      }
    }
    if (dead) { ++st.kill; nextq = Q_EMIT; }
    else {
      ++st.sca;
      const uint32_t pk_lo = P.U(U_PKLO, slot), pk_hi = P.U(U_PKHI, slot), ev = P.U(U_EV, slot);
#ifdef MCB_PHILOX2
      uint4 b, bnext;      // interaction block of flight ev, flight block of ev+1
      philox_block2((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 1u, 2u * ev + 2u, pk_lo, pk_hi, r.call_index, b, bnext);
#else
      const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 1u, pk_lo, pk_hi, r.call_index);
      const uint4 bnext = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 2u, pk_lo, pk_hi, r.call_index);
#endif
      const double u = P.F(F_U, slot), v = P.F(F_V, slot), w = P.F(F_W, slot);
      double u1, v1, w1;
      if (GR && r.lscattering_method1) {
        const Scat1Out o = scatter_method1<BANK>(lambda, p_icell, b, POLA, S[0], S[1], S[2], S[3], u, v, w);
        u1 = o.u1; v1 = o.v1; w1 = o.w1; S[0] = o.S0; S[1] = o.S1; S[2] = o.S2; S[3] = o.S3;
      } else {
        const float rand = u01(b.x), rand2 = u01(b.y), rand3 = u01(b.z);
        int itheta; double cospsi;
        if (r.lmethod_aniso1) angle_diff_theta_pos<SM>(m, r.p_lambda_in, p_icell, rand, rand2, itheta, cospsi);
        else hg(t_gfac<SM>(m, p_icell, lambda), rand, itheta, cospsi);
        if (r.lisotropic) { itheta = 1; cospsi = (double)__fsub_rn(__fmul_rn(2.0f, rand), 1.0f); }
        double sp, cp;
        mcb_sincospi((double)__fsub_rn(__fmul_rn(2.0f, rand3), 1.0f), &sp, &cp);     // PHI = PI*(2.0*rand-1.0): fp32 inner
        cdapres(cospsi, sp, cp, u, v, w, u1, v1, w1);
        if (POLA && r.lmethod_aniso1) scatter_stokes<BANK>(lambda, itheta, rand2, p_icell, S, u, v, w, u1, v1, w1);
      }
      misc |= MISC_SCATT;                                    // flag_scatt
      P.F(F_U, slot) = u1; P.F(F_V, slot) = v1; P.F(F_W, slot) = w1;
      if ((!TH && r.lmono) || POLA) { P.F(F_S0, slot) = S[0]; if (POLA) { QUV(0, slot) = S[1]; QUV(1, slot) = S[2]; QUV(2, slot) = S[3]; } }
      P.U(U_EV, slot) = ev + 1u;
      start_flight<BANK>(P, slot, P.F(F_PX, slot), P.F(F_PY, slot), P.F(F_PZ, slot), u1, v1, w1, bnext, misc);
      if (GR && r.capt_full) { POS0(0, slot) = P.F(F_PX, slot); POS0(1, slot) = P.F(F_PY, slot); POS0(2, slot) = P.F(F_PZ, slot); POS0(3, slot) = (double)idx; }
      nextq = Q_FLY;
      P.U(U_MISC, slot) = misc;
    }
  }
  return nextq;
}

// =============================================================================
// ABSORB: immediate re-emission, LTE (dust_transfer.f90:1353-1402) + start of the next flight
// =============================================================================
template <class G, bool SM, int BANK, int VAR>
__device__ __noinline__ int phase_absorb(int slot, bool valid, Stats& st) {
  constexpr bool GR = VAR == VAR_EXTRAS; constexpr bool TH = VAR == VAR_THERMAL; (void)GR; (void)TH;
  const DevModel& m = c_m; const DevRun& r = c_r;
  const bool POLA = r.lsepar_pola != 0;
  const Pool P = make_pool<SM, BANK>();
  using CellT = typename G::CellT;
  const unsigned lane = threadIdx.x & 31;
  int nextq = Q_NONE;
  double e_nRE = 0.0;
  if (valid) {
    const bool variable_dust = !SM && m.p_n_cells != 1;      // staged tables imply p_n_cells == 1
    uint32_t misc = P.U(U_MISC, slot);
    CellT cell; unpack_cell(P.U(U_C0A, slot), P.U(U_C0B, slot), cell);
    const int idx = tally_index(m, cell);
    const int p_icell = (variable_dust && idx >= 0) ? idx + 1 : 1;
    ++st.abs_;
    // idx < 0: interaction in a virtual cell (only with an inconsistent dark-zone mask): the packet is dropped below
    const LtePre pre = lte_prefetch(m, idx < 0 ? 0 : idx);      // in flight while the Philox blocks are computed
    const uint32_t pk_lo = P.U(U_PKLO, slot), pk_hi = P.U(U_PKHI, slot), ev = P.U(U_EV, slot);
#ifdef MCB_PHILOX2
    uint4 b, bnext;      // interaction block of flight ev, flight block of ev+1
    philox_block2((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 1u, 2u * ev + 2u, pk_lo, pk_hi, r.call_index, b, bnext);
#else
    const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 1u, pk_lo, pk_hi, r.call_index);
    const uint4 bnext = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 2u, pk_lo, pk_hi, r.call_index);
#endif
    int lambda;
    if (idx < 0) lambda = 0;
    else if (!GR || (r.lonly_LTE && !r.low_mem_th)) {
      // b.x is rand1: drawn but unused in the high-memory LTE branch (thermal_emission.f90:739-765)
      lambda = im_reemission_LTE<SM>(m, r, idx, p_icell, u01(b.y), pre);
    } else if (r.lonly_LTE) {
      lambda = im_reemission_LTE_lowmem<SM, BANK>(idx, p_icell, misc_lambda(misc), u01(b.x), u01(b.y), pre);
    } else {
      // the grain-regime draw (dust_transfer.f90:1379) has its own Philox block, so rand / rand2 / the
      // direction draws keep the words they have in the lonly_LTE case
      float sel = 0.0f;
      if (!r.lonly_nLTE) sel = u01(philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), (2u * ev + 1u) | 0x80000000u, pk_lo, pk_hi, r.call_index).x);
      const AbsOut o = absorb_grain_regimes<SM, BANK>(idx, p_icell, misc_lambda(misc), b, sel, P.F(F_S0, slot));
      lambda = o.lambda; e_nRE = o.e_nRE;
      if (r.lnRE) P.F(F_S0, slot) = o.S0;
    }
    if (lambda == 0) { ++st.kill; nextq = Q_EMIT; }      // Stokes(1) < tiny_real: packet dropped (dust_transfer.f90:1361-1364)
    else {
      double u, v, w;
      random_isotropic_direction(u01(b.z), u01(b.w), u, v, w);
      if (POLA) { QUV(0, slot) = 0.0; QUV(1, slot) = 0.0; QUV(2, slot) = 0.0; }
      misc = pack_misc(lambda, false, false, false, 0, misc_n_in_cell(misc));      // flag_star = flag_scatt = flag_ISM = .false.
      P.F(F_U, slot) = u; P.F(F_V, slot) = v; P.F(F_W, slot) = w;
      P.U(U_EV, slot) = ev + 1u;
      start_flight<BANK>(P, slot, P.F(F_PX, slot), P.F(F_PY, slot), P.F(F_PZ, slot), u, v, w, bnext, misc);
      if (GR && r.capt_full) { POS0(0, slot) = P.F(F_PX, slot); POS0(1, slot) = P.F(F_PY, slot); POS0(2, slot) = P.F(F_PZ, slot); POS0(3, slot) = (double)idx; }
      nextq = Q_FLY;
      P.U(U_MISC, slot) = misc;
    }
  }
  if (GR && r.lnRE) {      // E_abs_nRE (omp reduction in the reference, dust_transfer.f90:489): one atomic per warp
    for (int o = 16; o > 0; o >>= 1) e_nRE += __shfl_down_sync(0xffffffffu, e_nRE, o);
    if (lane == 0 && e_nRE != 0.0) atomicAdd(m.tally + m.lay.E_abs_nRE, e_nRE);
  }
  return nextq;
}

// =============================================================================
// ADOPT (straggler launch): a free slot takes the next packet the main launch parked
// =============================================================================
template <bool SM, int BANK>
__device__ __noinline__ int phase_adopt(int slot, bool valid) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  const Pool P = make_pool<SM, BANK>();
  const unsigned lane = threadIdx.x & 31;
  const unsigned need = __ballot_sync(0xffffffffu, valid);
  if (!need) return Q_NONE;
  unsigned long long* park_count = m.work + (12 + 2 * r.n_photons_loop);
  unsigned long long* park_head = m.work + (14 + 2 * r.n_photons_loop);
  const int leader = __ffs(need) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(park_head, (unsigned long long)__popc(need));
  base = __shfl_sync(0xffffffffu, base, leader);
  const unsigned long long j = base + __popc(need & ((1u << lane) - 1u));
  if (!valid || j >= __ldcg(park_count)) return Q_NONE;
  const double* rec = m.park + j * PARK_REC;
#pragma unroll
  for (int f = 0; f < 11; ++f) P.F(f, slot) = __ldcg(rec + f);
  const uint32_t* ru = reinterpret_cast<const uint32_t*>(rec + 11);
#pragma unroll
  for (int f = 0; f < NU32; ++f) P.U(f, slot) = __ldcg(ru + f);
  if (r.lsepar_pola) { QUV(0, slot) = __ldcg(rec + 16); QUV(1, slot) = __ldcg(rec + 17); QUV(2, slot) = __ldcg(rec + 18); }
  return (int)__ldcg(ru + NU32);
}

#ifdef MCB_DRAIN_PROBE
// development probe: when does the number of live packets of a block fall below 512, 256, ... 1, 0 after the
// packet counter ran dry?  max and sum over blocks of (t - t_dry) land in work[18+2n ..] (ns)
__device__ __forceinline__ void drain_probe(unsigned long long* work, int npl, unsigned live, int& k) {
  while (k < 11 && live <= (k < 10 ? (512u >> k) : 0u)) {
    const unsigned long long dry = __ldcg(work + 1);
    const unsigned long long now = globaltimer_ns();
    const unsigned long long dt = (dry != ~0ull && now > dry) ? now - dry : 0ull;
    atomicMax(work + (18 + 2 * npl + k), dt);
    atomicAdd(work + (29 + 2 * npl + k), dt);
    ++k;
  }
}
#endif

// =============================================================================
// The persistent photon-loop kernel: rounds of (claim a single-phase chunk -> run the phase -> regroup)
// =============================================================================
template <class G, bool SM, int BANK, int VAR>
__global__ void __launch_bounds__(MC_BLOCK, 1)
mc_photon_loop_kernel(const int adopt) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  const unsigned lane = threadIdx.x & 31;
  if (SM) stage_tables(m, r.p_lambda_in);
  const Pool P = make_pool<SM, BANK>();
  const bool POLA_ = r.lsepar_pola != 0;
  unsigned long long* park_count = m.work + (12 + 2 * r.n_photons_loop);
  if (threadIdx.x < 16) P.ctl[threadIdx.x] = 0;
  // every slot starts in the EMIT queue (main launch: emits new packets; straggler launch: adopts parked ones)
  for (int i = threadIdx.x; i < NQ * NP; i += MC_BLOCK) P.q[i] = (i < NP) ? (unsigned short)i : (unsigned short)0xFFFFu;   // Q_EMIT == 0
  __syncthreads();
  if (threadIdx.x == 0) { P.ctl[NQ + Q_EMIT] = NP; P.ctl[CTL_LIVE] = NP; }
  __syncthreads();
  const bool park_ok = r.park_enable && !adopt;
  // preferred queue order of this warp, one nibble per rank (see the scheduling loop); EMIT stays first everywhere:
  // free slots are refilled at once
  const unsigned q_order = ((threadIdx.x >> 5) & 3u) == 0u ? (unsigned)(Q_EMIT | (Q_ABS << 4) | (Q_SCAT << 8) | (Q_FLY << 12))
                         : ((threadIdx.x >> 5) & 3u) == 1u ? (unsigned)(Q_EMIT | (Q_SCAT << 4) | (Q_ABS << 8) | (Q_FLY << 12))
                                                           : (unsigned)(Q_EMIT | (Q_FLY << 4) | (Q_ABS << 8) | (Q_SCAT << 12));
  Stats st = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  SchedStats ss = {{0, 0, 0, 0}, {0, 0, 0, 0}};
#ifdef MCB_DRAIN_PROBE
  int probe_k = 0;
#endif
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(m.work + ((adopt ? 16 : 2) + 2 * r.n_photons_loop), globaltimer_ns());
  // ---- asynchronous scheduling: every warp repeatedly claims up to 32 entries of ONE queue (so all its
  // lanes run the same phase), preferring full chunks; partial chunks are only taken when no other warp
  // is busy (nothing more will arrive).  No block-wide barriers after this point.
  for (;;) {
    int qi = -1; unsigned h = 0, n = 0;
    if (lane == 0) {
      // prefer a full 32-entry chunk.  Partial chunks are taken at once when the pool is draining out
      // (few live packets: nothing will fill up), otherwise only after ~2 us without a full chunk.
      for (int polls = 0;; ++polls) {
#ifdef MCB_DEV
        if (c_r.debug_abort_dry && __ldcg(c_m.work + 1) != ~0ull) break;      // profiling aid: steady-state only
#endif
        if (park_ok && P.PARK()) break;                                       // the block is handing its packets over
        int best = -1; unsigned best_n = 0;
        const unsigned live = P.LIVE();
        unsigned emit_allow = 0xffffffffu;
        if (!adopt && c_r.inflight_frac_per_block > 0.0f) {
          // concurrency window: packets in flight <= max(floor, fraction of the packets sent so far) (see DevRun)
          const unsigned cap = max(c_r.inflight_floor, (unsigned)fminf(c_r.inflight_frac_per_block * (float)P.SENT(), (float)NP));
          const unsigned in_flight = live - (P.TAIL(Q_EMIT) - P.HEAD(Q_EMIT));
          emit_allow = cap > in_flight ? cap - in_flight : 0u;
        }
        // Queue order by SM sub-partition (warp id mod 4): two sub-partitions look at FLY first, one at ABSORB, one at
        // SCATTER, and each falls back to the other phases only when its own has no full chunk.  The phases are ~10-30 KB
        // of code each; warps that stay in one phase keep hitting the instruction cache of their sub-partition
        // (`no_instruction` was the top stall with every warp hopping between all phases).
#ifndef MCB_NO_PHASE_AFFINITY
        const unsigned order = q_order;
#else
        const unsigned order = 0x3210u;
#endif
#pragma unroll
        for (int kk = 0; kk < NQ; ++kk) {
          const int k = (int)((order >> (4 * kk)) & 15u);
          unsigned av = P.TAIL(k) - P.HEAD(k);
          if (k == Q_EMIT && av > emit_allow) av = emit_allow;
          if (av >= 32u) { best = k; best_n = 32u; break; }
          if (av > best_n) { best = k; best_n = av; }
        }
#ifdef MCB_DRAIN_PROBE
        if (threadIdx.x == 0 && P.DRYF()) drain_probe(c_m.work, c_r.n_photons_loop, live, probe_k);
#endif
        if (park_ok && live <= (unsigned)r.park_live && live > 0u && P.DRYF()) { P.PARK() = 1u; break; }
        // after the packet counter ran dry nothing refills the queues: waiting for full chunks only adds latency to the
        // long packet chains the call is now waiting for
        const bool dry = P.DRYF() != 0u;
        if (best >= 0 && (best_n >= (unsigned)c_r.min_chunk || live <= (dry ? (unsigned)c_r.drain_live_dry : DRAIN_LIVE) || polls >= (dry ? c_r.patience_dry : c_r.patience))) {
          const unsigned hh = P.HEAD(best);
          unsigned av = P.TAIL(best) - hh;
          if (best == Q_EMIT && av > emit_allow) av = emit_allow;
          const unsigned take = av < 32u ? av : 32u;
          if (take > 0 && atomicCAS((unsigned*)&P.ctl[best], hh, hh + take) == hh) { qi = best; h = hh; n = take; break; }
          continue;
        }
        if (best_n == 0u && live == 0u) break;       // every packet of this block is done
        __nanosleep(dry ? 60 : 250);
      }
    }
    qi = __shfl_sync(0xffffffffu, qi, 0);
    if (qi < 0) break;
    h = __shfl_sync(0xffffffffu, h, 0);
    n = __shfl_sync(0xffffffffu, n, 0);
    const bool valid = lane < n;
    int slot = 0;
    if (valid) {
      volatile unsigned short* e = P.Q(qi) + ((h + lane) & (NP - 1));
      unsigned short sv;
      while ((sv = *e) == 0xFFFFu) { }      // the producer reserved this entry and is about to write it
      *e = 0xFFFFu;
      slot = sv;
    }
    __threadfence_block();                   // acquire: packet state written before the entry
    // Run the phase; then regroup.  If (nearly) all lanes of this warp continue with the same next phase,
    // or the pool is draining out (small chunk), the warp keeps those packets and runs their next phase
    // directly; everything else goes back to the shared queues.
    bool mine = valid;
    for (;;) {
      { const unsigned mm = __ballot_sync(0xffffffffu, mine); if (lane == 0) { ss.visits[qi] += 1; ss.lanes[qi] += __popc(mm); } }
      int nextq;
      switch (qi) {
        case Q_EMIT: nextq = adopt ? phase_adopt<SM, BANK>(slot, mine) : phase_emit<G, SM, BANK, VAR>(slot, mine, st); break;
        case Q_ABS:  nextq = phase_absorb<G, SM, BANK, VAR>(slot, mine, st); break;
        case Q_SCAT: nextq = phase_scatter<G, SM, BANK, VAR>(slot, mine, st); break;
        default:     nextq = phase_fly<G, SM, BANK, VAR>(slot, mine, st); break;
      }
      if (!mine) nextq = Q_NONE;
      int keep = -1, keep_n = 0, total_n = 0;
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const int c = __popc(__ballot_sync(0xffffffffu, nextq == k));
        total_n += c;
        if (c > keep_n) { keep_n = c; keep = k; }
      }
      unsigned live_now = 0;
      if (lane == 0) live_now = (park_ok && P.PARK()) ? 0xFFFFFFFFu : P.LIVE();      // parking: everything goes back to the queues
#ifdef MCB_DRAIN_PROBE
      if (threadIdx.x == 0 && P.DRYF() && live_now != 0xFFFFFFFFu) drain_probe(c_m.work, c_r.n_photons_loop, live_now, probe_k);
#endif
      live_now = __shfl_sync(0xffffffffu, live_now, 0);                      // warp-uniform decision
      const bool cont = keep_n > 0 && live_now != 0xFFFFFFFFu && (keep_n >= 28 || live_now <= (P.DRYF() ? (unsigned)c_r.drain_live_dry : DRAIN_LIVE));
      const int pushq = (cont && nextq == keep) ? Q_NONE + 1 : nextq;      // kept lanes are not pushed
      push_next(P, slot, pushq, mine, lane);
      if (!cont) break;
      mine = mine && (nextq == keep);
      qi = keep;
      __threadfence_block();
    }
  }

  // ---- straggler hand-over: every warp has left the loop, so every live packet sits in a queue ----
  if (park_ok) {
    __syncthreads();
    if (P.PARK()) {
      for (int k = Q_ABS; k < NQ; ++k) {          // EMIT entries are free slots (the counter is dry)
        const unsigned hq = P.HEAD(k), nq = P.TAIL(k) - hq;
        for (unsigned t = threadIdx.x; t < nq; t += MC_BLOCK) {
          const int slot = P.Q(k)[(hq + t) & (NP - 1)];
          const unsigned long long j = atomicAdd(park_count, 1ull);
          double* rec = m.park + j * PARK_REC;
#pragma unroll
          for (int f = 0; f < 11; ++f) rec[f] = P.F(f, slot);
          uint32_t* ru = reinterpret_cast<uint32_t*>(rec + 11);
#pragma unroll
          for (int f = 0; f < NU32; ++f) ru[f] = P.U(f, slot);
          ru[NU32] = (uint32_t)k;
          if (POLA_) { rec[16] = QUV(0, slot); rec[17] = QUV(1, slot); rec[18] = QUV(2, slot); }
        }
      }
    }
  }

  // ---- diagnostics (not part of the reference) ----
  double* stt = m.tally + m.lay.stats;
  auto flush = [&](int k, unsigned long long vv) {
    for (int o = 16; o > 0; o >>= 1) vv += __shfl_down_sync(0xffffffffu, vv, o);
    if (lane == 0 && vv) atomicAdd(stt + k, (double)vv);
  };
  flush(STAT_PACKETS, st.pk); flush(STAT_STEPS, st.steps); flush(STAT_INTERACT, st.inter); flush(STAT_SCATT, st.sca);
  flush(STAT_ABS, st.abs_); flush(STAT_KILLED, st.kill); flush(STAT_ESCAPED, st.esc); flush(STAT_BOUNCE, st.bounce);
  flush(STAT_MRW_WALKS, st.mrw_w); flush(STAT_MRW_STEPS, st.mrw_s);
  if (lane == 0) {
    unsigned long long* dbg = m.work + (4 + 2 * r.n_photons_loop);
    for (int k = 0; k < 4; ++k) { atomicAdd(dbg + k, (unsigned long long)ss.visits[k]); atomicAdd(dbg + 4 + k, (unsigned long long)ss.lanes[k]); }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(m.work + (3 + 2 * r.n_photons_loop), globaltimer_ns());
  // last block to leave: end of the main launch / of the straggler launch (diagnostics)
  if (threadIdx.x == 0) atomicMax(m.work + ((adopt ? 15 : 13) + 2 * r.n_photons_loop), (unsigned long long)globaltimer_ns());
}

}  // namespace mcb
