// The Monte Carlo photon-loop kernel translation unit.  Compiled WITHOUT --fmad=false:
// the packet loop is statistical (atomics make the summation order non-deterministic anyway),
// so FMA contraction is allowed here, while the deterministic sub-kernels in api.cu keep
// the reference's non-contracted arithmetic and stay bit-exact against the oracle.
#include <cstdlib>
#include "handle.cuh"
#include "transport.cuh"

using namespace mcb;

// End-of-main-launch events of the last two calls that hand stragglers over (any handle).  A new main launch waits
// for the one before the previous: at most ONE main launch is pending while another runs, so the SMs the
// running one leaves free go to straggler launches (high-priority stream), not to a third call's blocks.
// (__constant__ banks and SMs are per device, so both pieces of bookkeeping are kept per device)
constexpr int MCB_MAX_DEV = 16;
static cudaEvent_t g_main_hist_dev[MCB_MAX_DEV][2] = {};
// last user of each constant bank (every kernel variant of a bank reads the same __constant__ copy)
struct BankGuard { const mcb_handle* owner = nullptr; cudaEvent_t done = nullptr; };
static BankGuard bank_guard_dev[MCB_MAX_DEV][MCB_BANKS];
void mcb_forget_handle(const mcb_handle* h) {      // called by finalize after the handle's streams were synchronised
  for (auto& e : g_main_hist_dev[h->device % MCB_MAX_DEV]) if (e == h->ev_main) e = nullptr;
  for (auto& g : bank_guard_dev[h->device % MCB_MAX_DEV]) if (g.owner == h) g.owner = nullptr;
}


template <class G, bool SM, int BANK, int VAR>
static int launch_bank(mcb_handle* h, const DevRun& dr) {
  const size_t smem = (SM ? (size_t)h->m.sm.total_words * 8 : 0) + pool_bytes(dr.lsepar_pola != 0);
  auto kern = mc_photon_loop_kernel<G, SM, BANK, VAR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // model + run parameters -> constant memory, ordered on the handle's stream
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
  // the previous launch of this handle must be over before its constant bank is rewritten; so must the last
  // launch of any OTHER handle that maps to the same bank (more handles than banks)
  CK(cudaStreamSynchronize(h->stream));
  {
    BankGuard& g = bank_guard_dev[h->device % MCB_MAX_DEV][BANK];
    if (g.owner && g.owner != h && g.done) CK(cudaEventSynchronize(g.done));
    if (!g.done) CK(cudaEventCreateWithFlags(&g.done, cudaEventDisableTiming));
  }
  const int bank = BANK;
  CK(cudaMemcpyToSymbolAsync(c_mm, &h->m, sizeof(DevModel), (size_t)bank * sizeof(DevModel), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyToSymbolAsync(c_rr, &dr, sizeof(DevRun), (size_t)bank * sizeof(DevRun), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));          // dr / h->m are host stack / heap values
  int blocks = h->n_sm;                          // persistent: one 512-thread block (1024 packets in flight) per SM
  const int n2 = dr.park_enable ? h->straggler_sms : 0;    // blocks of the straggler launch (mcfost_b200_set_overlap)
  if (dr.park_enable) blocks -= h->overlap_sms;            // SMs the main launch leaves to straggler launches
  // test knob: fewer blocks = fewer packets in flight.  Immediate re-emission reads RUNNING tallies, so a
  // run whose packet budget is not >> 1024 x blocks sees them at a different stage than a 16-thread CPU run.
  { const char* e = getenv("MCB_BLOCKS"); if (e && atoi(e) > 0 && atoi(e) < blocks) blocks = atoi(e); }
  cudaEvent_t* g_main_hist = g_main_hist_dev[h->device % MCB_MAX_DEV];
  if (n2 > 0 && g_main_hist[0] && g_main_hist[0] != h->ev_main) CK(cudaStreamWaitEvent(h->stream, g_main_hist[0], 0));
  CK(cudaEventRecord(h->ev0, h->stream));
  kern<<<blocks, MC_BLOCK, smem, h->stream>>>(0);
  CK(cudaGetLastError());
  if (n2 > 0) {      // the stragglers the main launch parked, on the SMs the main launches leave free.  Highest
    // stream priority: when SMs free up, these few blocks go before the pending main blocks of other handles.
    CK(cudaEventRecord(h->ev_main, h->stream));
    g_main_hist[0] = g_main_hist[1]; g_main_hist[1] = h->ev_main;
    CK(cudaStreamWaitEvent(h->stream_hi, h->ev_main, 0));
    kern<<<n2, MC_BLOCK, smem, h->stream_hi>>>(1);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev_strag, h->stream_hi));
    CK(cudaStreamWaitEvent(h->stream, h->ev_strag, 0));      // everything later on the handle's stream is ordered after it
  }
  CK(cudaEventRecord(h->ev1, h->stream));
  { BankGuard& g = bank_guard_dev[h->device % MCB_MAX_DEV][BANK]; g.owner = h; CK(cudaEventRecord(g.done, h->stream)); }
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
      case GK_CYL3D: return launch_grains<GeomCyl<true, false>>(h, dr);
      case GK_SPH2D: return launch_grains<GeomSph<false, false>>(h, dr);
      case GK_SPH3D: return launch_grains<GeomSph<true, false>>(h, dr);
      case GK_VOR:   return launch_grains<GeomVor>(h, dr);
    }
  }
  switch (h->gk) {
    case GK_CYL2D: return sm ? launch_one<GeomCyl<false, true>, true>(h, dr) : launch_one<GeomCyl<false, false>, false>(h, dr);
    case GK_CYL3D: return sm ? launch_one<GeomCyl<true, true>, true>(h, dr) : launch_one<GeomCyl<true, false>, false>(h, dr);
    case GK_SPH2D: return sm ? launch_one<GeomSph<false, true>, true>(h, dr) : launch_one<GeomSph<false, false>, false>(h, dr);
    case GK_SPH3D: return sm ? launch_one<GeomSph<true, true>, true>(h, dr) : launch_one<GeomSph<true, false>, false>(h, dr);
    case GK_VOR:   return launch_one<GeomVor, false>(h, dr);
  }
  return MCB_ERR_BAD_ARG;
}
