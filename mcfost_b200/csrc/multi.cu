// Multi-GPU behind the C ABI: one host process drives the n GPUs of a node, so that a Fortran
// caller gets all of them from ONE call.  Packets shard by chunk (rank g takes the chunks c with (c-1) mod n == g, the
// Philox streams depend on the packet id only, so the union of the ranks' packets is the single-GPU set); grid and
// tables are replicated; after the kernels ONE group of NCCL all-reduces over NVLink merges every tally the call
// produced:   sum  packed fp64 block (xKJ_abs, xJ_abs, n_phot_envoyes, sed x 9, stats, E_abs_nRE), xI_scatt, I_spec,
//                  I_spec_star, photon maps, origin tallies, xN_abs      (thermal_emission.f90:668, output.f90:3084-3102,
//                                                                          dust_ray_tracing.f90:661,689)
//             min  xT_ech            (Temp_LTE with id = 0 starts from minval(xT_ech(icell,:)), thermal_emission.f90:683)
//             max  xT_ech_1grain, xT_ech_1grain_nRE      (maxval, thermal_emission.f90:823,977)
// NCCL is loaded at run time (dlopen of libnccl.so.2) the first time a multi-GPU object with n > 1 is created, so the
// library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <vector>
#include "handle.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef enum { ncclSuccess = 0 } ncclResult_t;
// values of nccl.h (stable since NCCL 2.0)
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2, ncclMin = 3 };
struct Nccl {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return false;
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};
Nccl g_nccl;
}  // namespace

struct mcb_multi {
  std::vector<mcb_handle*> h;
  std::vector<ncclComm_t> comm;
  char err[512] = {0};
};

static int mfail(mcb_multi* m, int code, const char* msg) { snprintf(m->err, sizeof m->err, "%s", msg); return code; }

extern "C" {

const char* mcfost_b200_multi_last_error(const mcb_multi* m) { return m ? m->err : "null multi-GPU object"; }

int mcfost_b200_multi_init(int n_gpus, const int* devices, mcb_multi** out) {
  if (!out || n_gpus < 1) return MCB_ERR_BAD_ARG;
  *out = nullptr;
  mcb_multi* m = new mcb_multi();
  std::vector<int> dev((size_t)n_gpus);
  for (int i = 0; i < n_gpus; ++i) dev[i] = devices ? devices[i] : i;
  for (int i = 0; i < n_gpus; ++i) {
    mcb_handle* hh = nullptr;
    const int rc = mcfost_b200_init(dev[i], &hh);
    if (rc) { for (auto* q : m->h) mcfost_b200_finalize(q); delete m; return rc; }
    m->h.push_back(hh);
  }
  if (n_gpus > 1) {
    if (!g_nccl.load()) { for (auto* q : m->h) mcfost_b200_finalize(q); delete m; return MCB_ERR_UNSUPPORTED; }
    m->comm.resize((size_t)n_gpus);
    if (g_nccl.CommInitAll(m->comm.data(), n_gpus, dev.data()) != ncclSuccess) { for (auto* q : m->h) mcfost_b200_finalize(q); delete m; return MCB_ERR_CUDA; }
  }
  *out = m;
  return MCB_OK;
}

void mcfost_b200_multi_finalize(mcb_multi* m) {
  if (!m) return;
  for (auto c : m->comm) if (c) g_nccl.CommDestroy(c);
  for (auto* q : m->h) mcfost_b200_finalize(q);
  delete m;
}

int mcfost_b200_multi_n_gpus(const mcb_multi* m) { return m ? (int)m->h.size() : 0; }
mcb_handle* mcfost_b200_multi_handle(mcb_multi* m, int i) { return (m && i >= 0 && i < (int)m->h.size()) ? m->h[(size_t)i] : nullptr; }

#define MULTI_EACH(call)                                                                                         \
  if (!m) return MCB_ERR_BAD_ARG;                                                                                \
  for (auto* hh : m->h) { const int rc = (call); if (rc) { snprintf(m->err, sizeof m->err, "%s", mcfost_b200_last_error(hh)); return rc; } } \
  return MCB_OK;
int mcfost_b200_multi_upload_grid(mcb_multi* m, const mcb_grid* g) { MULTI_EACH(mcfost_b200_upload_grid(hh, g)) }
int mcfost_b200_multi_upload_dark_zone(mcb_multi* m, const int32_t* dz) { MULTI_EACH(mcfost_b200_upload_dark_zone(hh, dz)) }
int mcfost_b200_multi_upload_opacity(mcb_multi* m, const mcb_opacity* o) { MULTI_EACH(mcfost_b200_upload_opacity(hh, o)) }
int mcfost_b200_multi_upload_emission(mcb_multi* m, const mcb_emission* e) { MULTI_EACH(mcfost_b200_upload_emission(hh, e)) }
int mcfost_b200_multi_upload_grains(mcb_multi* m, const mcb_grains* g) { MULTI_EACH(mcfost_b200_upload_grains(hh, g)) }

// The drop-in for mc_photon_loop on n GPUs: blocking.  r->rank / r->n_ranks are overwritten (rank g = GPU g).
int mcfost_b200_multi_run(mcb_multi* m, const mcb_run_params* r, mcb_tallies* out) {
  if (!m || !r) return MCB_ERR_BAD_ARG;
  const int n = (int)m->h.size();
  for (int g = 0; g < n; ++g) {
    mcb_run_params rg = *r;
    rg.rank = g; rg.n_ranks = n;
    const int rc = mcfost_b200_launch(m->h[(size_t)g], &rg);
    if (rc) { snprintf(m->err, sizeof m->err, "GPU %d: %s", g, mcfost_b200_last_error(m->h[(size_t)g])); return rc; }
  }
  if (n > 1) {
    bool ok = true;
    auto red = [&](auto get_ptr, size_t count, int dtype, int op) {
      if (!count) return;
      for (int g = 0; g < n; ++g) {
        mcb_handle* hh = m->h[(size_t)g];
        void* p = get_ptr(hh);
        cudaSetDevice(hh->device);
        if (g_nccl.AllReduce(p, p, count, dtype, op, m->comm[(size_t)g], hh->stream) != ncclSuccess) ok = false;
      }
    };
    mcb_handle* h0 = m->h[0];
    g_nccl.GroupStart();
    red([](mcb_handle* q) { return (void*)q->m.tally; }, (size_t)h0->n_tally, ncclFloat64, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.xT_ech; }, (size_t)h0->m.n_cells, ncclInt32, ncclMin);
    red([](mcb_handle* q) { return (void*)q->m.xI; }, (size_t)h0->n_xI, ncclFloat32, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.I_spec; }, (size_t)h0->n_Ispec, ncclFloat32, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.I_spec_star; }, h0->n_Ispec ? (size_t)h0->m.n_cells : 0, ncclFloat32, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.smap; }, (size_t)h0->n_map, ncclFloat64, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.star_origin; }, (size_t)h0->n_org, ncclFloat64, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.xN; }, (size_t)h0->n_xN, ncclFloat64, ncclSum);
    red([](mcb_handle* q) { return (void*)q->m.gr.xT_1g; }, (size_t)h0->n_1g, ncclInt32, ncclMax);
    red([](mcb_handle* q) { return (void*)q->m.gr.xT_1g_nRE; }, (size_t)h0->n_1g_nRE, ncclInt32, ncclMax);
    if (g_nccl.GroupEnd() != ncclSuccess) ok = false;
    if (!ok) return mfail(m, MCB_ERR_CUDA, "ncclAllReduce failed");
  }
  for (int g = 0; g < n; ++g) {
    const int rc = mcfost_b200_sync(m->h[(size_t)g]);
    if (rc) { snprintf(m->err, sizeof m->err, "GPU %d: %s", g, mcfost_b200_last_error(m->h[(size_t)g])); return rc; }
  }
  if (out) {
    const int rc = mcfost_b200_download(m->h[0], r, out);
    if (rc) { snprintf(m->err, sizeof m->err, "%s", mcfost_b200_last_error(m->h[0])); return rc; }
  }
  return MCB_OK;
}

// Temp_finale on the merged tallies (every GPU holds them after the all-reduce): GPU 0 answers
int mcfost_b200_multi_temp_finale(mcb_multi* m, float* Tdust) {
  if (!m) return MCB_ERR_BAD_ARG;
  const int rc = mcfost_b200_temp_finale(m->h[0], Tdust);
  if (rc) snprintf(m->err, sizeof m->err, "%s", mcfost_b200_last_error(m->h[0]));
  return rc;
}

}  // extern "C"
