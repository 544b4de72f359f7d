// The packet-per-warp kernel of the thermal step (LTE re-emission, scattering method 2): the low-latency companion of
// mc_photon_loop_kernel.
//
// Why it exists.  A packet's events are a dependent chain, and the packet-per-lane kernel (transport.cuh) needs ~13 000
// cycles per flight + interaction of ONE packet (measured: a single packet alone on the GPU, profiles/r02_latency.md):
// ~2500 sequential instructions, most of them Philox rounds, sincospi / log / sqrt polynomials and bisections.  Whenever
// few packets are in flight that latency, not throughput, is what the call waits for: the last packets of every call (a
// packet trapped in the optically thick inner rim makes 1e5 absorb / re-emit cycles), and the whole of a call with the
// reference's own packet budget (1.28e5 packets; the packets in flight are capped to a fraction of those already sent).
//
// How.  One warp runs one packet, every lane holding the same packet state in registers, and the work that does not
// depend on the packet's state is done for the next 16 events at once, one event per lane:
//   * lanes 0-15 compute the flight blocks (Philox block 2e: tau, interaction-type draw), lanes 16-31 the interaction
//     blocks (2e+1) of events e .. e+15, and from them the isotropic re-emission direction, the scattering angle (s11
//     bisection / HG) and the azimuth -- every sincospi, logf and bisection of 16 events in one SIMD pass;
//   * the event loop then only does what depends on the packet: wall distances, the optical-depth test, the tally,
//     the temperature of the cell and the new wavelength, the latter as ONE 32-lane probe of the CDF (a ballot) instead
//     of a 6-step bisection.
// Mutable global memory (the running tallies) is read by lane 0 and broadcast, so the lanes stay bit-identical; only
// lane 0 deposits.  Same Philox streams as the packet-per-lane kernel (a packet may start in one kernel and finish in
// the other: the straggler hand-over of transport.cuh parks packets that this kernel adopts).  Directions of the
// look-ahead are kept in fp32 and renormalised in fp64 (the thermal step is compared statistically).
#pragma once
#include "transport.cuh"

namespace mcb {

#ifndef MCB_ENG_BLOCK
#define MCB_ENG_BLOCK 512
#endif
constexpr int ENG_BLOCK = MCB_ENG_BLOCK;          // 16 warps per SM, <= 128 registers per thread
constexpr int ENG_LOOK = 16;            // events of look-ahead

// what one lane of the look-ahead holds for event (base + lane % 16)
struct Look {
  float tau, ralb;                      // lanes 0-15: flight block
  float rand2;                          // lanes 16-31: interaction block: wavelength draw / Mueller interpolation
  float iu, iv, iw;                     // isotropic direction (absorption)
  float cpsi, sphi, cphi;               // scattering: cos(psi), sin / cos(phi)
  int itheta;
};

template <bool SM, int BANK>
__device__ __forceinline__ Look look_ahead(uint32_t pk_lo, uint32_t pk_hi, uint32_t base, unsigned lane) {
  const DevModel& m = c_m; const DevRun& r = c_r;
  Look L;
  const uint32_t e = base + (lane & 15u);
  const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * e + (lane >> 4), pk_lo, pk_hi, r.call_index);
  L.tau = tau_of_rand(u01(b.x)); L.ralb = u01(b.y);
  L.rand2 = u01(b.y);
  // absorption: isotropic direction from (z, w) (random_isotropic_direction)
  {
    const float wz = 2.0f * u01(b.z) - 1.0f;
    const float uv = sqrtf(fmaxf(0.0f, 1.0f - wz * wz));
    float sp, cp; sincospif(2.0f * u01(b.w) - 1.0f, &sp, &cp);
    L.iu = uv * cp; L.iv = uv * sp; L.iw = wz;
  }
  // scattering, method 2 (dust_transfer.f90:1318-1348): (x, y) -> angle, z -> azimuth.  The HG branch needs g(lambda):
  // it is evaluated in the event loop (cpsi is then recomputed there)
  {
    const float rand = u01(b.x), rand2 = u01(b.y), rand3 = u01(b.z);
    int itheta = 1; double cospsi = 0.0;
    if (r.lmethod_aniso1 && m.p_n_cells == 1) angle_diff_theta_pos<SM>(m, r.p_lambda_in, 1, rand, rand2, itheta, cospsi);
    if (r.lisotropic) { itheta = 1; cospsi = (double)(2.0f * rand - 1.0f); }
    L.cpsi = (float)cospsi; L.itheta = itheta;
    float sp, cp; sincospif(2.0f * rand3 - 1.0f, &sp, &cp);
    L.sphi = sp; L.cphi = cp;
  }
  return L;
}

template <class G, bool SM, int BANK>
__global__ void __launch_bounds__(ENG_BLOCK, 1)
mc_warp_engine_kernel(const unsigned long long emit_limit, const int adopt, const int sm_kf, const int sm_kdB) {
  // sm_kf / sm_kdB: word offsets of kf_dark(n_cells) / kdB_dT_CDF(n_lambda, n_T) staged in shared memory by this kernel
  // (it has no packet pool, so the space is free), or -1: a lone warp cannot hide an L2 round trip per cell crossing
  constexpr int VAR = VAR_THERMAL;
  const DevModel& m = c_m; const DevRun& r = c_r;
  using CellT = typename G::CellT;
  using Hit = typename G::Hit;
  const unsigned lane = threadIdx.x & 31;
  const bool POLA = r.lsepar_pola != 0;
  const bool variable_dust = !SM && m.p_n_cells != 1;
  if (SM) stage_tables(m, r.p_lambda_in);
  {
    double* sd = reinterpret_cast<double*>(mcb_smem_raw);
    if (sm_kf >= 0) for (int i = threadIdx.x; i < m.n_cells; i += blockDim.x) sd[sm_kf + i] = m.kf_dark[i];
    if (sm_kdB >= 0) for (int i = threadIdx.x; i < m.n_lambda * m.n_T; i += blockDim.x) sd[sm_kdB + i] = m.kdB[i];
    __syncthreads();
  }
  auto kf_of = [&](int idx) -> double { return sm_kf >= 0 ? smd_ld(m, sm_kf + idx) : __ldg(m.kf_dark + idx); };
  auto kdB_of = [&](int l, int t, int p_icell) -> double {      // l, t 1-based
    return sm_kdB >= 0 ? smd_ld(m, sm_kdB + (l - 1) + m.n_lambda * (t - 1)) : t_kdB<SM>(m, l, t, p_icell); };
  const int b0 = 2 + 2 * r.n_photons_loop;
  unsigned long long* park_count = m.work + (b0 + 10);
  unsigned long long* park_head = m.work + (b0 + 12);
  unsigned long long* emit_counter = m.work + (b0 + 15);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicExch(m.work + (b0 + (adopt ? 14 : 40)), globaltimer_ns());
  unsigned n_pk = 0, n_steps = 0, n_inter = 0, n_sca = 0, n_abs = 0, n_kill = 0, n_esc = 0, n_bounce = 0, n_mrw_w = 0, n_mrw_s = 0;   // lane 0 counts

  for (;;) {
    // ------------------------------------------------------------------ next packet of this warp
    double x0, y0, z0, u, v, w, S0, extr, xo, yo, zo, Sq = 0.0, Su = 0.0, Sv = 0.0;
    CellT c0, c_old, c_start; null_cell(c_old);
    uint32_t pk_lo, pk_hi, ev, misc;
    float ralb;
    int entry = Q_FLY;                       // parked packets may be waiting for their interaction (Q_ABS / Q_SCAT)
    if (adopt) {
      unsigned long long j = 0;
      if (lane == 0) j = atomicAdd(park_head, 1ull);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= __ldcg(park_count)) break;
      const double* rec = m.park + j * PARK_REC;
      x0 = __ldcg(rec + F_PX); y0 = __ldcg(rec + F_PY); z0 = __ldcg(rec + F_PZ);
      xo = __ldcg(rec + F_OX); yo = __ldcg(rec + F_OY); zo = __ldcg(rec + F_OZ);
      u = __ldcg(rec + F_U); v = __ldcg(rec + F_V); w = __ldcg(rec + F_W);
      S0 = __ldcg(rec + F_S0); extr = __ldcg(rec + F_EXTR);
      const uint32_t* ru = reinterpret_cast<const uint32_t*>(rec + 11);
      unpack_cell(__ldcg(ru + U_C0A), __ldcg(ru + U_C0B), c0);
      unpack_cell(__ldcg(ru + U_COA), __ldcg(ru + U_COB), c_old);
      pk_lo = __ldcg(ru + U_PKLO); pk_hi = __ldcg(ru + U_PKHI); ev = __ldcg(ru + U_EV); misc = __ldcg(ru + U_MISC);
      ralb = __uint_as_float(__ldcg(ru + U_RALB));
      entry = (int)__ldcg(ru + NU32);
      if (POLA) { Sq = __ldcg(rec + 16); Su = __ldcg(rec + 17); Sv = __ldcg(rec + 18); }
      c_start = c0;      // n_iteractions_in_cell restarts with the hand-over (the packet-per-lane kernel does not keep it)
    } else {
      unsigned long long g = 0;
      if (lane == 0) g = atomicAdd(emit_counter, 1ull);
      g = __shfl_sync(0xffffffffu, g, 0);
      if (g >= emit_limit) break;
      const int first_local = r.nnfot1_start + ((r.rank - ((r.nnfot1_start - 1) % r.n_ranks) + r.n_ranks) % r.n_ranks);
      const unsigned long long lc = g / r.n_per_chunk, idx_in_chunk = g % r.n_per_chunk;
      const unsigned long long packet = ((unsigned long long)(first_local + (int)lc * r.n_ranks - 1) << 40) + idx_in_chunk;
      pk_lo = (uint32_t)packet; pk_hi = (uint32_t)(packet >> 32);
      ++n_pk;
      const Emitted<CellT> e = emit_packet_core<G, SM, BANK, VAR>(pk_lo, pk_hi);
      if (lane == 0) atomicAdd(m.tally + m.lay.n_env + (e.lambda - 1), 1.0);
      if (!e.lintersect) {       // the packet never enters the model: straight to the detector (dust_transfer.f90:545-552)
        if (!e.flag_ISM) {
          const double S[4] = {e.S0, 0.0, 0.0, 0.0};
          if (lane == 0) capteur<BANK>(e.lambda, e.u, e.v, e.w, S, e.flag_star, false);
          ++n_esc;
        }
        continue;
      }
      x0 = e.x; y0 = e.y; z0 = e.z; u = e.u; v = e.v; w = e.w; S0 = e.S0; c0 = e.cell; c_start = c0;
      xo = x0; yo = y0; zo = z0;
      ev = 1u;
      misc = pack_misc(e.lambda, e.flag_star, false, e.flag_ISM, 0, 0);
      const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u, pk_lo, pk_hi, r.call_index);
      extr = (double)tau_of_rand(u01(b.x)); ralb = u01(b.y);
      misc = misc_set_istar(misc, intersect_stars(m, x0, y0, z0, u, v, w));
    }
    int lambda = misc_lambda(misc);
    int n_in_cell = misc_n_in_cell(misc);
    int idx_c = tally_index(m, c0);          // tally index of c0 (-1: virtual cell), kept up to date wherever c0 changes
    uint32_t la_base = 0u; bool la_have = false;
    int pre_idx = -1; LtePre pre_raw; pre_raw.xkj = 0.0; pre_raw.vol = 1.0; pre_raw.Ti = 2; double dep_own = 0.0;
    Look L; L.tau = 0; L.ralb = 0; L.rand2 = 0; L.iu = 0; L.iv = 0; L.iw = 1; L.cpsi = 1; L.sphi = 0; L.cphi = 1; L.itheta = 1;

    // ------------------------------------------------------------------ the packet's life
    for (;;) {
      bool finished = false;
      // ---- flight ev (physical_length, optical_depth.f90:77-178); skipped for a parked packet that waits for its interaction
      if (entry == Q_FLY) {
        const DirInv dinv = dir_invariants(u, v, w);
        const int i_star_hit = misc_istar(misc);
        CellT c_star; null_cell(c_star);
        if (i_star_hit > 0) cell_of_id(m, m.star_icell[i_star_hit - 1], c_star);      // the cell of the star this flight points at
        // running tally / temperature index / volume of the cell the flight starts in, requested now and used by the
        // absorption if the flight ends in the same cell (the rule for a trapped packet): the L2 round trip overlaps the flight
        pre_idx = idx_c;
        if (pre_idx >= 0) pre_raw = lte_prefetch(m, pre_idx);
        dep_own = 0.0;
        // opacities of the flight's wavelength (constant dust: one shared-memory read per flight instead of two per cell)
        const double kap_l = variable_dust ? 0.0 : t_kappa<SM>(m, 1, lambda);
        const double kabs_S0 = variable_dust ? 0.0 : t_kappa_abs<SM>(m, 1, lambda) * S0;
        for (;;) {
          if (idx_c < 0 && G::test_exit(m, c0, x0, y0, z0)) {      // (a real cell is never an exit)
            if (!misc_ism(misc)) {
              const double S[4] = {S0, Sq, Su, Sv};
              if (lane == 0) capteur<BANK>(lambda, u, v, w, S, misc_star(misc), misc_scatt(misc));
              ++n_esc;
            }
            finished = true; break;
          }
          if (i_star_hit > 0 && same_cell(c0, c_star)) { ++n_kill; finished = true; break; }
          const int idx = idx_c;
          double opacity = 0.0, kf = 0.0;
          int p_icell = 1;
          if (idx >= 0) {
            p_icell = variable_dust ? idx + 1 : 1;
            kf = kf_of(idx);
          }
          if (idx >= 0) {
            if (signbit(kf)) {      // dark-zone bounce (optical_depth.f90:104-112)
              u = -u; v = -v; w = -w;
              c0 = c_old; x0 = xo; y0 = yo; z0 = zo; idx_c = tally_index(m, c0);
              ++n_bounce;
              break;
            }
            opacity = (variable_dust ? t_kappa<SM>(m, p_icell, lambda) : kap_l) * kf;
          }
          ++n_steps;
          const Hit h = G::distance(m, dinv, x0, y0, z0, u, v, w, c0, c_old);
          double l_contrib = hit_l_contrib(h), l = h.l;
          const double tau_c = l_contrib * opacity;
          bool lstop = false;
          if (tau_c > extr) { lstop = true; l_contrib = l_contrib * mc_div(extr, tau_c); l = hit_l_void(h) + l_contrib; }
          else extr = extr - tau_c;
          if (idx >= 0) {      // save_radiation_field (radiation_field.f90:53-54)
            const double dep = variable_dust ? t_kappa_abs<SM>(m, p_icell, lambda) * l_contrib * S0 : kabs_S0 * l_contrib;
            if (idx == pre_idx) dep_own += dep;
            if (lane == 0) {
              atomicAdd(m.tally + m.lay.xKJ + idx, dep);
              if (r.lxJ) atomicAdd(m.tally + m.lay.xJ + idx + (size_t)m.n_cells * (lambda - 1), l_contrib * S0);
            }
          }
          if (lstop) {
            x0 = x0 + l * u; y0 = y0 + l * v; z0 = z0 + l * w;
            if (!G::is_vor && m.l3D && m.kind == 1) { c0 = G::index(m, x0, y0, z0); idx_c = tally_index(m, c0); }
            break;
          }
          double x1, y1, z1; CellT c1;
          G::advance(m, h, x0, y0, z0, u, v, w, c0, x1, y1, z1, c1);
          xo = x0; yo = y0; zo = z0; c_old = c0;
          x0 = x1; y0 = y1; z0 = z1; c0 = c1; idx_c = tally_index(m, c1);
        }
        if (finished) break;
        ++n_inter;
        n_in_cell = same_cell(c0, c_start) ? min(n_in_cell + 1, 255) : 0;      // dust_transfer.f90:1242-1249
      }
      // ---- interaction ending flight ev (dust_transfer.f90:1260-1402)
      const int idx = idx_c;
      const int p_icell = (variable_dust && idx >= 0) ? idx + 1 : 1;
      if (idx < 0) { ++n_kill; break; }      // interaction in a virtual cell (inconsistent dark-zone mask): drop the packet
      {
      if (!la_have || ev - la_base >= (uint32_t)ENG_LOOK) { la_base = ev; la_have = true; L = look_ahead<SM, BANK>(pk_lo, pk_hi, la_base, lane); }
      const int k = (int)(ev - la_base);
      const bool scatter = (entry == Q_FLY) ? (ralb < t_albedo<SM>(m, p_icell, lambda)) : (entry == Q_SCAT);
      entry = Q_FLY;
      if (scatter) {
        ++n_sca;
        double cospsi = (double)__shfl_sync(0xffffffffu, L.cpsi, 16 + k);
        const double sp = (double)__shfl_sync(0xffffffffu, L.sphi, 16 + k), cp = (double)__shfl_sync(0xffffffffu, L.cphi, 16 + k);
        int itheta = __shfl_sync(0xffffffffu, L.itheta, 16 + k);
        const float rand2 = __shfl_sync(0xffffffffu, L.rand2, 16 + k);
        if (!r.lisotropic && (!r.lmethod_aniso1 || variable_dust)) {
          // the angle depends on the packet (Henyey-Greenstein with g(lambda), or a cell-dependent phase function):
          // not in the look-ahead; the interaction block's first two draws are needed here
          const uint4 b = philox_block((uint32_t)r.seed, (uint32_t)(r.seed >> 32), 2u * ev + 1u, pk_lo, pk_hi, r.call_index);
          if (r.lmethod_aniso1) angle_diff_theta_pos<SM>(m, r.p_lambda_in, p_icell, u01(b.x), u01(b.y), itheta, cospsi);
          else hg(t_gfac<SM>(m, p_icell, lambda), u01(b.x), itheta, cospsi);
        }
        // sin / cos(phi) come from fp32: put them back on the unit circle
        const double nphi = rsqrt(sp * sp + cp * cp);
        double u1, v1, w1;
        cdapres(cospsi, sp * nphi, cp * nphi, u, v, w, u1, v1, w1);
        if (POLA && r.lmethod_aniso1) {
          double S[4] = {S0, Sq, Su, Sv};
          scatter_stokes<BANK>(lambda, itheta, rand2, p_icell, S, u, v, w, u1, v1, w1);
          S0 = S[0]; Sq = S[1]; Su = S[2]; Sv = S[3];
        }
        u = u1; v = v1; w = w1;
        misc |= MISC_SCATT;
      } else {
        ++n_abs;
        // Temp_LTE + im_reemission_LTE (thermal_emission.f90:649-771): running tally read by lane 0
        LtePre pre;
        if (idx == pre_idx) {      // requested at the start of the flight; this flight's own deposits in the cell are added
          pre.xkj = __shfl_sync(0xffffffffu, pre_raw.xkj, 0) + dep_own; pre.Ti = __shfl_sync(0xffffffffu, pre_raw.Ti, 0); pre.vol = pre_raw.vol;
        } else pre = lte_prefetch_w<true>(m, idx);
        pre_idx = -1;
        int Ti; double frac_T2;
        temp_lte<SM>(m, r, idx, p_icell, pre, Ti, frac_T2);
        const double frac_T1 = 1.0 - frac_T2;
        const float rand2 = __shfl_sync(0xffffffffu, L.rand2, 16 + k);
        // the bisection of :753-765 returns the first l in 1..n_lambda-1 with rand2 <= proba(l), else n_lambda: one probe per lane
        int lam = m.n_lambda;
        for (int base = 0; base < m.n_lambda - 1; base += 32) {
          const int l = base + (int)lane + 1;
          bool ok = false;
          if (l <= m.n_lambda - 1) {
            const double proba = frac_T1 * kdB_of(l, Ti - 1, p_icell) + frac_T2 * kdB_of(l, Ti, p_icell);
            ok = !((double)rand2 > proba);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, ok);
          if (bal) { lam = base + __ffs(bal); break; }
        }
        lambda = lam;
        const double iu = (double)__shfl_sync(0xffffffffu, L.iu, 16 + k), iv = (double)__shfl_sync(0xffffffffu, L.iv, 16 + k), iw = (double)__shfl_sync(0xffffffffu, L.iw, 16 + k);
        const double nrm = rsqrt(iu * iu + iv * iv + iw * iw);
        u = iu * nrm; v = iv * nrm; w = iw * nrm;
        Sq = 0.0; Su = 0.0; Sv = 0.0;
        misc = pack_misc(lambda, false, false, false, 0, 0);
      }
      ++ev;
      }
      // ---- modified random walk (dust_transfer.f90:1222-1239)
      bool try_mrw = false;
      if (r.lMRW && n_in_cell > 5) try_mrw = __shfl_sync(0xffffffffu, (int)mrw_worth_trying<G, BANK>(c0, idx, x0, y0, z0), 0) != 0;
      if (try_mrw) {
        const MrwOut o = mrw_walk<G, SM, BANK, true>(c0, idx, p_icell, x0, y0, z0, S0, pk_lo, pk_hi, ev);
        if (o.steps) {
          x0 = o.x; y0 = o.y; z0 = o.z; u = o.u; v = o.v; w = o.w; lambda = o.lambda; ev = o.ev;
          Sq = 0.0; Su = 0.0; Sv = 0.0;
          misc = pack_misc(lambda, false, false, false, 0, 0);
          ++n_mrw_w; n_mrw_s += o.steps;
        }
      }
      // ---- start flight ev: tau and the interaction-type draw of block 2 ev
      if (ev - la_base >= (uint32_t)ENG_LOOK) { la_base = ev; L = look_ahead<SM, BANK>(pk_lo, pk_hi, la_base, lane); }      // (unsigned difference: also after an MRW jump)
      {
        const int k2 = (int)(ev - la_base);
        extr = (double)__shfl_sync(0xffffffffu, L.tau, k2);
        ralb = __shfl_sync(0xffffffffu, L.ralb, k2);
      }
      xo = x0; yo = y0; zo = z0; null_cell(c_old); c_start = c0;
      misc = misc_set_istar(misc, intersect_stars(m, x0, y0, z0, u, v, w));
    }
  }

  // ---- diagnostics (lane 0 of every warp counted its packets' events)
  if (lane == 0) {
    double* stt = m.tally + m.lay.stats;
    auto add = [&](int k, unsigned vv) { if (vv) atomicAdd(stt + k, (double)vv); };
    add(STAT_PACKETS, n_pk); add(STAT_STEPS, n_steps); add(STAT_INTERACT, n_inter); add(STAT_SCATT, n_sca); add(STAT_ABS, n_abs);
    add(STAT_KILLED, n_kill); add(STAT_ESCAPED, n_esc); add(STAT_BOUNCE, n_bounce); add(STAT_MRW_WALKS, n_mrw_w); add(STAT_MRW_STEPS, n_mrw_s);
  }
  if (threadIdx.x == 0) atomicMax(m.work + (b0 + (adopt ? 13 : 41)), (unsigned long long)globaltimer_ns());
}

}  // namespace mcb
