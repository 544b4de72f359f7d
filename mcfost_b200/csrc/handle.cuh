// Shared between api.cu (host API + deterministic kernels, --fmad=false) and
// mc_kernel.cu (the Monte Carlo photon-loop kernel, FMA contraction allowed).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/mcfost_b200.h"
#include "model.cuh"

struct mcb_handle {
  int device = 0;
  int bank = 0;                           // constant-memory bank of this handle's launches (see transport.cuh)
  int n_sm = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream_hi = nullptr;       // highest priority: straggler launches are dispatched before pending main blocks of other handles
  cudaEvent_t ev_main = nullptr, ev_strag = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  mcb::DevModel m;
  mcb::GridKind gk = mcb::GK_CYL2D;
  bool has_grid = false, has_op = false, has_em = false, has_gr = false, launched = false;
  mcb_grains gr_host{};                   // which optional grain tables were supplied (pointers are not dereferenced after upload)
  int64_t n_xN = 0;                       // xN_abs size of the last launch
  int64_t n_map = 0, n_org = 0;           // photon-map / origin tally sizes of the last launch
  int64_t n_1g = 0, n_1g_nRE = 0;         // extents of xT_ech_1grain / xT_ech_1grain_nRE of the last launch
  std::map<std::string, void*> bufs;      // named device allocations
  std::map<std::string, size_t> buf_bytes;
  int64_t n_tally = 0, n_xI = 0, n_Ispec = 0;
  bool lay_xJ = false;
  int lay_nsed = -1;
  int n_photons_loop_alloc = 0;
  int n_type_flux = 1;
  int eps_ntf = 0, eps_lambda = 0;                     // eps_dust1 on the device: N_type_flux and wavelength it was built for
  int rt1_n_rt = 0, rt1_pola = 0, rt1_contrib = 0;      // shape of the xI_scatt tally of the last rt1 launch (init_dust_source_fct1)
  int straggler_sms = 0;                  // (kept for the set_overlap signature)
  int launches_last_call = 0;             // kernels of this library launched by the last mcfost_b200_launch
  int overlap_sms = 0;                    // mcfost_b200_set_overlap: SMs reserved for straggler launches (0 = off)
  std::vector<double> host_kappa_factor;  // host copies used to build kf_dark (kappa_factor | dark flag)
  std::vector<uint8_t> host_dark;
  bool kf_dark_stale = true;
  bool em_on_device = false;              // prob_E_cell / frac_E_stars / frac_E_disk were built by mcfost_b200_repartition_energie
  bool mrw_ready = false;                 // zeta table + mean opacities of the modified random walk are on the device
  char err[512] = {0};
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      snprintf(h->err, sizeof h->err, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      return MCB_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

static int fail(mcb_handle* h, int code, const char* msg) {
  snprintf(h->err, sizeof h->err, "%s", msg);
  return code;
}

// (re)allocate a named device buffer and optionally fill it from host memory
template <class T>
static int put(mcb_handle* h, const char* name, const T* src, size_t n, const T** dst) {
  *dst = nullptr;
  if (!src || n == 0) return MCB_OK;
  size_t bytes = n * sizeof(T);
  void*& p = h->bufs[name];
  if (p && h->buf_bytes[name] != bytes) { cudaFree(p); p = nullptr; }
  if (!p) { CK(cudaMalloc(&p, bytes)); h->buf_bytes[name] = bytes; }
  CK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyHostToDevice, h->stream));
  *dst = (const T*)p;
  return MCB_OK;
}
template <class T>
static int reserve(mcb_handle* h, const char* name, size_t n, T** dst) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  void*& p = h->bufs[name];
  if (p && h->buf_bytes[name] != bytes) { cudaFree(p); p = nullptr; }
  if (!p) { CK(cudaMalloc(&p, bytes)); h->buf_bytes[name] = bytes; }
  *dst = (T*)p;
  return MCB_OK;
}


// defined in mc_kernel.cu
int mcb_launch_mc(mcb_handle* h, const mcb::DevRun& dr);
void mcb_forget_handle(const mcb_handle* h);
