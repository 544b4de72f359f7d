// The Monte Carlo photon-loop kernel translation unit.  Compiled WITHOUT --fmad=false:
// the packet loop is statistical (atomics make the summation order non-deterministic anyway),
// so FMA contraction is allowed here, while the deterministic sub-kernels in api.cu keep
// the reference's non-contracted arithmetic and stay bit-exact against the oracle.
#include "handle.cuh"
#include "transport.cuh"

using namespace mcb;

template <class G, bool SM>
static int launch_one(mcb_handle* h, const DevRun& dr) {
  const size_t smem = SM ? (size_t)h->m.sm.total_words * 8 : 0;
  auto kern = mc_photon_loop_kernel<G, SM>;
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MC_BLOCK, smem));
  if (per_sm < 1) per_sm = 1;
  const int blocks = h->n_sm * per_sm;          // persistent: exactly one resident wave over the 148 SMs
  CK(cudaEventRecord(h->ev0, h->stream));
  kern<<<blocks, MC_BLOCK, smem, h->stream>>>(h->m, dr);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, h->stream));
  return MCB_OK;
}

int mcb_launch_mc(mcb_handle* h, const DevRun& dr) {
  const bool sm = h->m.sm.enabled != 0;
  switch (h->gk) {
    case GK_CYL2D: return sm ? launch_one<GeomCyl<false, true>, true>(h, dr) : launch_one<GeomCyl<false, false>, false>(h, dr);
    case GK_CYL3D: return sm ? launch_one<GeomCyl<true, true>, true>(h, dr) : launch_one<GeomCyl<true, false>, false>(h, dr);
    case GK_SPH2D: return sm ? launch_one<GeomSph<false, true>, true>(h, dr) : launch_one<GeomSph<false, false>, false>(h, dr);
    case GK_SPH3D: return sm ? launch_one<GeomSph<true, true>, true>(h, dr) : launch_one<GeomSph<true, false>, false>(h, dr);
    case GK_VOR:   return launch_one<GeomVor, false>(h, dr);
  }
  return MCB_ERR_BAD_ARG;
}
