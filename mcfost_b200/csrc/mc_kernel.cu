// The Monte Carlo photon-loop kernel translation unit.  Compiled WITHOUT --fmad=false:
// the packet loop is statistical (atomics make the summation order non-deterministic anyway),
// so FMA contraction is allowed here, while the deterministic sub-kernels in api.cu keep
// the reference's non-contracted arithmetic and stay bit-exact against the oracle.
#include <cmath>
#include <cstdlib>
#include <mutex>
#define MCB_MC_FAST_MATH 1      // see geom_rz.cuh (mc_rcp / mc_div): the Monte Carlo kernels only
#include "handle.cuh"
#include "transport.cuh"
#include "warp_engine.cuh"

using namespace mcb;

// End-of-main-launch events of the last two calls that hand stragglers over (any handle).  A new main launch waits
// for the one before the previous: at most ONE main launch is pending while another runs, so the SMs the
// running one leaves free go to straggler launches (high-priority stream), not to a third call's blocks.
// (__constant__ banks and SMs are per device, so both pieces of bookkeeping are kept per device; a mutex guards them,
// several host threads may launch on their own handles)
constexpr int MCB_MAX_DEV = 64;
static std::mutex g_launch_mutex;
static cudaEvent_t g_main_hist_dev[MCB_MAX_DEV][2] = {};
// last user of each constant bank (every kernel variant of a bank reads the same __constant__ copy)
struct BankGuard { const mcb_handle* owner = nullptr; cudaEvent_t done = nullptr; };
static BankGuard bank_guard_dev[MCB_MAX_DEV][MCB_BANKS];
void mcb_forget_handle(const mcb_handle* h) {      // called by finalize after the handle's streams were synchronised
  std::lock_guard<std::mutex> lock(g_launch_mutex);
  for (auto& e : g_main_hist_dev[h->device % MCB_MAX_DEV]) if (e == h->ev_main) e = nullptr;
  for (auto& g : bank_guard_dev[h->device % MCB_MAX_DEV]) if (g.owner == h) g.owner = nullptr;
}

__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }

// Thermal step: which kernel sends the packets.
// The packet-per-lane kernel keeps up to 1024 packets per SM in flight and is throughput-oriented: ~5 us per event of
// one packet.  A call whose whole budget is small is latency-bound from start to end (the longest packet chain decides)
// and runs on the packet-per-warp kernel alone (~0.5 us per event, 16 packets in flight per SM); larger calls run on the
// packet-per-lane kernel, whose scheduler ramps the packets in flight up with the packets sent (DevRun::inflight_*),
// and only their last packets go to the packet-per-warp kernel.
static unsigned long long engine_first_packets(unsigned long long n_total, int blocks, double frac) {
  const unsigned long long small = (unsigned long long)std::ceil(4.0 * 128.0 * blocks / frac);      // 1.2e6 packets on 148 SMs at 1/16
  return n_total <= small ? n_total : 0ull;
}

template <class G, bool SM, int BANK, int VAR>
static int launch_bank(mcb_handle* h, DevRun dr) {
  if (h->device < 0 || h->device >= MCB_MAX_DEV) return fail(h, MCB_ERR_UNSUPPORTED, "device index >= 64");
  std::lock_guard<std::mutex> lock(g_launch_mutex);
  const size_t smem_tables = SM ? (size_t)h->m.sm.total_words * 8 : 0;
  const size_t smem = smem_tables + pool_bytes(dr.lsepar_pola != 0);
  auto kern = mc_photon_loop_kernel<G, SM, BANK, VAR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks_full = h->n_sm - h->overlap_sms;            // persistent: one block (1024 packets in flight) per SM; set_overlap leaves SMs to the tail kernels of other handles
  // ---- which kernels run this call
  constexpr bool TH = VAR == VAR_THERMAL;
  const double frac = dr.inflight_frac_per_block;            // (api.cu stores the caller's fraction here; divided by the block count below)
  unsigned long long n_engine_first = 0;                     // thermal step: packets sent by the packet-per-warp kernel
  bool lane_kernel = true;
  int blocks = blocks_full;
  if (TH) {
    n_engine_first = engine_first_packets(dr.n_packets_total, blocks_full, frac);
    lane_kernel = n_engine_first < dr.n_packets_total;
    dr.park_enable = 1;                                      // the last packets of the per-lane kernel are finished by the packet-per-warp kernel
    dr.inflight_floor = 32u;
    dr.inflight_frac_per_block = (float)(frac / blocks);
  } else {
    dr.inflight_frac_per_block = 0.0f;                       // no cap inside the kernel
    if (dr.letape_th && dr.count_sent) {
      // other thermal modes (per-grain branches): keep the packets in flight a small fraction of the budget by using fewer blocks
      const unsigned long long want = (unsigned long long)(frac * (double)dr.n_packets_total / NP);
      blocks = (int)std::min<unsigned long long>((unsigned long long)blocks_full, std::max<unsigned long long>(1ull, want));
    }
    dr.park_enable = 0;
  }
  if (dr.lsepar_pola) {        // Stokes Q,U,V slabs (one per block), L2-resident
    void*& q = h->bufs["quv"];
    const size_t bytes = (size_t)h->n_sm * 3 * NP * sizeof(double);
    if (q && h->buf_bytes["quv"] != bytes) { cudaFree(q); q = nullptr; }
    if (!q) { CK(cudaMalloc(&q, bytes)); h->buf_bytes["quv"] = bytes; }
    h->m.quv = (double*)q;
  }
  if (dr.park_enable) {
    void*& pk = h->bufs["park"];
    const size_t bytes = (size_t)h->n_sm * PARK_LIVE * PARK_REC * sizeof(double);
    if (pk && h->buf_bytes["park"] != bytes) { cudaFree(pk); pk = nullptr; }
    if (!pk) { CK(cudaMalloc(&pk, bytes)); h->buf_bytes["park"] = bytes; }
    h->m.park = (double*)pk;
  }
  if (dr.lMRW) {
    void*& lr = h->bufs["mrw_lR"];
    const size_t bytes_lr = (size_t)h->m.n_cells * 2 * sizeof(float);
    if (lr && h->buf_bytes["mrw_lR"] != bytes_lr) { cudaFree(lr); lr = nullptr; }
    if (!lr) { CK(cudaMalloc(&lr, bytes_lr)); h->buf_bytes["mrw_lR"] = bytes_lr; }
    CK(cudaMemsetAsync(lr, 0, bytes_lr, h->stream));      // (a hint only; reset per call so that a call does not depend on the previous one)
    h->m.mrw_lR = (float*)lr;
  }
  // the previous launch of this handle must be over before its constant bank is rewritten; so must the last
  // launch of any OTHER handle that maps to the same bank (more handles than banks)
  CK(cudaStreamSynchronize(h->stream));
  {
    BankGuard& g = bank_guard_dev[h->device][BANK];
    if (g.owner && g.owner != h && g.done) CK(cudaEventSynchronize(g.done));
    if (!g.done) CK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
  }
  const int bank = BANK;
  CK(cudaMemcpyToSymbolAsync(c_mm, &h->m, sizeof(DevModel), (size_t)bank * sizeof(DevModel), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyToSymbolAsync(c_rr, &dr, sizeof(DevRun), (size_t)bank * sizeof(DevRun), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));          // dr / h->m are host stack / heap values
  cudaEvent_t* g_main_hist = g_main_hist_dev[h->device];
  const bool overlap = h->overlap_sms > 0 && dr.park_enable;
  if (overlap && g_main_hist[0] && g_main_hist[0] != h->ev_main) CK(cudaStreamWaitEvent(h->stream, g_main_hist[0], 0));
  CK(cudaEventRecord(h->ev0, h->stream));
  h->launches_last_call = 0;
  if constexpr (TH) {
    auto eng = mc_warp_engine_kernel<G, SM, BANK>;
    // the packet-per-warp kernel has no packet pool: kf_dark(n_cells) and, with constant dust, kdB_dT_CDF(n_lambda, n_T)
    // are staged in the shared memory that frees, when they fit
    int sm_kf = -1, sm_kdB = -1;
    size_t smem_eng = smem_tables;
    if (smem_eng + (size_t)h->m.n_cells * 8 <= 160 * 1024) { sm_kf = (int)(smem_eng / 8); smem_eng += (size_t)h->m.n_cells * 8; }
    if (h->m.p_n_cells == 1 && smem_eng + (size_t)h->m.n_lambda * h->m.n_T * 8 <= 200 * 1024) { sm_kdB = (int)(smem_eng / 8); smem_eng += (size_t)h->m.n_lambda * h->m.n_T * 8; }
    if (smem_eng > 48 * 1024) CK(cudaFuncSetAttribute(eng, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_eng));
    if (n_engine_first > 0) {
      eng<<<h->n_sm, ENG_BLOCK, smem_eng, h->stream>>>(n_engine_first, 0, sm_kf, sm_kdB);
      CK(cudaGetLastError());
      ++h->launches_last_call;
    }
    if (lane_kernel) {
      set_u64_kernel<<<1, 1, 0, h->stream>>>(h->m.work, n_engine_first);      // the per-lane kernel continues the packet counter
      kern<<<blocks, MC_BLOCK, smem, h->stream>>>(0);
      CK(cudaGetLastError());
      h->launches_last_call += 2;
      // its stragglers, on the packet-per-warp kernel.  With set_overlap: highest stream priority, so that when SMs free
      // up these blocks go before the pending main blocks of other handles.
      cudaStream_t st2 = h->stream;
      if (overlap) {
        CK(cudaEventRecord(h->ev_main, h->stream));
        g_main_hist[0] = g_main_hist[1]; g_main_hist[1] = h->ev_main;
        CK(cudaStreamWaitEvent(h->stream_hi, h->ev_main, 0));
        st2 = h->stream_hi;
      }
      eng<<<h->n_sm, ENG_BLOCK, smem_eng, st2>>>(0ull, 1, sm_kf, sm_kdB);
      CK(cudaGetLastError());
      ++h->launches_last_call;
      if (overlap) {
        CK(cudaEventRecord(h->ev_strag, h->stream_hi));
        CK(cudaStreamWaitEvent(h->stream, h->ev_strag, 0));      // everything later on the handle's stream is ordered after it
      }
    }
  } else {
    kern<<<blocks, MC_BLOCK, smem, h->stream>>>(0);
    CK(cudaGetLastError());
    ++h->launches_last_call;
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  { BankGuard& g = bank_guard_dev[h->device][BANK]; g.owner = h; CK(cudaEventRecord(g.done, h->stream)); }
  return MCB_OK;
}

template <class G, bool SM>
static int launch_one(mcb_handle* h, const DevRun& dr) {
  // the thermal step (wavelength drawn per packet, chunks end on packets sent, no ray-tracing tallies) has its own variant
  const bool th = dr.letape_th && !dr.lmono && dr.count_sent && !dr.rt1 && !dr.rt2;
  switch (h->bank % MCB_BANKS) {
    case 0:  return th ? launch_bank<G, SM, 0, VAR_THERMAL>(h, dr) : launch_bank<G, SM, 0, VAR_GENERIC>(h, dr);
    case 1:  return th ? launch_bank<G, SM, 1, VAR_THERMAL>(h, dr) : launch_bank<G, SM, 1, VAR_GENERIC>(h, dr);
    default: return th ? launch_bank<G, SM, 2, VAR_THERMAL>(h, dr) : launch_bank<G, SM, 2, VAR_GENERIC>(h, dr);
  }
}
// per-grain modes (scattering method 1, nLTE / qRE re-emission): tables in global memory, GR = true kernels
template <class G>
static int launch_grains(mcb_handle* h, const DevRun& dr) {
  return (h->bank % 2) == 0 ? launch_bank<G, false, 0, VAR_EXTRAS>(h, dr) : launch_bank<G, false, 1, VAR_EXTRAS>(h, dr);      // two banks are enough here (guarded)
}

int mcb_launch_mc(mcb_handle* h, const DevRun& dr) {
  bool sm = h->m.sm.enabled != 0;
  // the pool (~127-151 KB) and the staged tables must fit the 227 KB of one SM
  if (sm && (size_t)h->m.sm.total_words * 8 + pool_bytes(dr.lsepar_pola != 0) > 227 * 1024) sm = false;
  if (dr.lscattering_method1 || !dr.lonly_LTE || dr.low_mem_th || dr.capt_full || dr.lspot || dr.lweight_emission || dr.lxN) {
    switch (h->gk) {
      case GK_CYL2D: return launch_grains<GeomCyl<false, false>>(h, dr);
#ifndef MCB_DEV_CYL2D_ONLY
      case GK_CYL3D: return launch_grains<GeomCyl<true, false>>(h, dr);
      case GK_SPH2D: return launch_grains<GeomSph<false, false>>(h, dr);
      case GK_SPH3D: return launch_grains<GeomSph<true, false>>(h, dr);
      case GK_VOR:   return launch_grains<GeomVor>(h, dr);
#endif
      default: return MCB_ERR_BAD_ARG;
    }
  }
  switch (h->gk) {
#ifndef MCB_DEV_CYL2D_ONLY      // development builds: only the kernels of the headline configuration (fast compile)
    case GK_CYL3D: return sm ? launch_one<GeomCyl<true, true>, true>(h, dr) : launch_one<GeomCyl<true, false>, false>(h, dr);
    case GK_SPH2D: return sm ? launch_one<GeomSph<false, true>, true>(h, dr) : launch_one<GeomSph<false, false>, false>(h, dr);
    case GK_SPH3D: return sm ? launch_one<GeomSph<true, true>, true>(h, dr) : launch_one<GeomSph<true, false>, false>(h, dr);
    case GK_VOR:   return launch_one<GeomVor, false>(h, dr);
#endif
    case GK_CYL2D: return sm ? launch_one<GeomCyl<false, true>, true>(h, dr) : launch_one<GeomCyl<false, false>, false>(h, dr);
    default: break;
  }
  return MCB_ERR_BAD_ARG;
}
