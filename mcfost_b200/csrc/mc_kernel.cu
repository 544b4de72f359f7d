// The Monte Carlo photon-loop kernel translation unit.  Compiled WITHOUT --fmad=false:
// the packet loop is statistical (atomics make the summation order non-deterministic anyway),
// so FMA contraction is allowed here, while the deterministic sub-kernels in api.cu keep
// the reference's non-contracted arithmetic and stay bit-exact against the oracle.
#include <cstdlib>
#include "handle.cuh"
#include "transport.cuh"

using namespace mcb;

template <class G, bool SM, int BANK, bool GR>
static int launch_bank(mcb_handle* h, const DevRun& dr) {
  const size_t smem = (SM ? (size_t)h->m.sm.total_words * 8 : 0) + pool_bytes(dr.lsepar_pola != 0);
  auto kern = mc_photon_loop_kernel<G, SM, BANK, GR>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // model + run parameters -> constant memory, ordered on the handle's stream
  if (dr.lsepar_pola) {        // Stokes Q,U,V slabs (one per block), L2-resident
    void*& q = h->bufs["quv"];
    const size_t bytes = (size_t)h->n_sm * 3 * NP * sizeof(double);
    if (q && h->buf_bytes["quv"] != bytes) { cudaFree(q); q = nullptr; }
    if (!q) { CK(cudaMalloc(&q, bytes)); h->buf_bytes["quv"] = bytes; }
    h->m.quv = (double*)q;
  }
  // the previous launch of this handle must be over before its constant bank is rewritten
  CK(cudaStreamSynchronize(h->stream));
  const int bank = BANK;
  CK(cudaMemcpyToSymbolAsync(c_mm, &h->m, sizeof(DevModel), (size_t)bank * sizeof(DevModel), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyToSymbolAsync(c_rr, &dr, sizeof(DevRun), (size_t)bank * sizeof(DevRun), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));          // dr / h->m are host stack / heap values
  int blocks = h->n_sm;                          // persistent: one 512-thread block (1024 packets in flight) per SM
  // test knob: fewer blocks = fewer packets in flight.  Immediate re-emission reads RUNNING tallies, so a
  // run whose packet budget is not >> 1024 x blocks sees them at a different stage than a 16-thread CPU run.
  { const char* e = getenv("MCB_BLOCKS"); if (e && atoi(e) > 0 && atoi(e) < blocks) blocks = atoi(e); }
  CK(cudaEventRecord(h->ev0, h->stream));
  kern<<<blocks, MC_BLOCK, smem, h->stream>>>();
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  return MCB_OK;
}

template <class G, bool SM>
static int launch_one(mcb_handle* h, const DevRun& dr) {
  return (h->bank % MCB_BANKS) == 0 ? launch_bank<G, SM, 0, false>(h, dr) : launch_bank<G, SM, 1, false>(h, dr);
}
// per-grain modes (scattering method 1, nLTE / qRE re-emission): tables in global memory, GR = true kernels
template <class G>
static int launch_grains(mcb_handle* h, const DevRun& dr) {
  return (h->bank % MCB_BANKS) == 0 ? launch_bank<G, false, 0, true>(h, dr) : launch_bank<G, false, 1, true>(h, dr);
}

int mcb_launch_mc(mcb_handle* h, const DevRun& dr) {
  bool sm = h->m.sm.enabled != 0;
  // the pool (~127-151 KB) and the staged tables must fit the 227 KB of one SM
  if (sm && (size_t)h->m.sm.total_words * 8 + pool_bytes(dr.lsepar_pola != 0) > 227 * 1024) sm = false;
  if (dr.lscattering_method1 || !dr.lonly_LTE) {
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
