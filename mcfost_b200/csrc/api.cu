// Host side of the C ABI declared in include/mcfost_b200.h: device memory
// ownership, uploads, kernel launches, tally download.  There is NO CPU
// fallback anywhere in this library: without a CUDA device every entry point
// fails with MCB_ERR_NO_DEVICE.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include <cub/cub.cuh>
#include "handle.cuh"
#include "cells.cuh"

using namespace mcb;

static char g_err[256] = {0};

extern "C" {

const char* mcfost_b200_last_error(const mcb_handle* h) { return h ? h->err : g_err; }

int mcfost_b200_init(int device, mcb_handle** out) {
  if (!out) return MCB_ERR_BAD_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    snprintf(g_err, sizeof g_err, "no CUDA device (%s); mcfost_b200 has no CPU fallback", cudaGetErrorString(e));
    return MCB_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { snprintf(g_err, sizeof g_err, "device %d out of range (0..%d)", device, n - 1); return MCB_ERR_BAD_ARG; }
  mcb_handle* h = new mcb_handle();
  memset(&h->m, 0, sizeof h->m);
  h->device = device;
  { static int next_bank = 0; h->bank = next_bank++; }      // handles share MCB_BANKS constant banks round-robin (guarded in mc_kernel.cu)
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return MCB_ERR_CUDA; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  h->n_sm = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return MCB_ERR_CUDA; }
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->stream_hi, cudaStreamNonBlocking, hi) != cudaSuccess) { delete h; return MCB_ERR_CUDA; }
    cudaEventCreateWithFlags(&h->ev_main, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_strag, cudaEventDisableTiming);
  }
  *out = h;
  return MCB_OK;
}

void mcfost_b200_finalize(mcb_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  for (auto& kv : h->bufs) if (kv.second) cudaFree(kv.second);
  cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1);
  cudaStreamSynchronize(h->stream_hi);
  mcb_forget_handle(h);
  cudaEventDestroy(h->ev_main); cudaEventDestroy(h->ev_strag);
  cudaStreamDestroy(h->stream_hi);
  cudaStreamDestroy(h->stream);
  delete h;
}

// Scheduling diagnostics of the last launch (not part of the reference interface):
// out[0] = ms from kernel start until the global packet counter ran dry (steady-state phase),
// out[1] = ms of the whole kernel by the device clock, out[2..5] = chunk visits per phase
// (EMIT, ABSORB, SCATTER, FLY), out[6..9] = valid lanes summed over those visits,
// out[10] = packets handed over to the straggler launch (mcfost_b200_set_overlap), out[11] / out[12] = ms until the
// last block of the main launch / of the straggler launch left.
int mcfost_b200_debug_counters(mcb_handle* h, double* out) {
  if (!h || !out) return MCB_ERR_BAD_ARG;
  if (!h->launched || !h->m.work) return fail(h, MCB_ERR_STATE, "no launch yet");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const int n = 48 + 2 * h->n_photons_loop_alloc;
  std::vector<unsigned long long> w((size_t)n);
  CK(cudaMemcpy(w.data(), h->m.work, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  const int b = 2 + 2 * h->n_photons_loop_alloc;
  const double t0 = (double)w[b], t1 = (double)w[b + 1];
  out[0] = (w[1] == ~0ull || w[1] == 0ull) ? -1.0 : ((double)w[1] - t0) * 1e-6;
  out[1] = (t1 - t0) * 1e-6;
  for (int k = 0; k < 8; ++k) out[2 + k] = (double)w[b + 2 + k];
  out[10] = (double)w[b + 10];
  out[11] = w[b + 11] ? ((double)w[b + 11] - t0) * 1e-6 : -1.0;      // ms until the last block of the main launch left
  out[12] = w[b + 13] ? ((double)w[b + 13] - t0) * 1e-6 : -1.0;      // ms until the straggler launch ended
  out[14] = w[b + 14] ? ((double)w[b + 14] - t0) * 1e-6 : -1.0;      // ms until the straggler launch started
  out[15] = (double)h->launches_last_call;      // kernels the last mcfost_b200_launch started (photon-loop kernels + the counter hand-over)
  out[13] = t0 * 1e-6;      // device globaltimer at the start of the main launch, ms (to line up calls on two handles)
  if (getenv("MCB_DRAIN_PROBE_PRINT")) {      // development builds (-DMCB_DRAIN_PROBE) only
    for (int k = 0; k < 11; ++k) fprintf(stderr, "drain probe: live <= %4d  max %8.2f ms  mean %8.2f ms after dry\n", k < 10 ? (512 >> k) : 0, (double)w[b + 16 + k] * 1e-6, (double)w[b + 27 + k] * 1e-6 / (double)h->n_sm);
  }
  return MCB_OK;
}

int mcfost_b200_stream(mcb_handle* h, uint64_t* s) { if (!h || !s) return MCB_ERR_BAD_ARG; *s = (uint64_t)(uintptr_t)h->stream; return MCB_OK; }

// ---------------------------------------------------------------------------
int mcfost_b200_upload_grid(mcb_handle* h, const mcb_grid* g) {
  if (!h || !g) return MCB_ERR_BAD_ARG;
  CK(cudaSetDevice(h->device));
  DevModel& m = h->m;
  if (g->kind < MCB_GRID_CYL || g->kind > MCB_GRID_VORONOI) return fail(h, MCB_ERR_BAD_ARG, "unknown grid kind");
  if (g->n_stars > MAX_STARS) return fail(h, MCB_ERR_UNSUPPORTED, "more than 15 stars (4 bits of the packed packet state)");
  // limits of the packed packet state (transport.cuh pack_cell: zj in 16 signed bits, k in 16 bits)
  if (g->kind != MCB_GRID_VORONOI && (g->nz > 32766 || g->n_az > 65535)) return fail(h, MCB_ERR_UNSUPPORTED, "nz > 32766 or n_az > 65535");
  m.kind = g->kind; m.l3D = g->l3D; m.n_rad = g->n_rad; m.nz = g->nz; m.n_az = g->n_az; m.n_cells = g->n_cells;
  m.nj = g->l3D ? 2 * g->nz : g->nz;
  m.Rmax2 = g->Rmax2; m.zmaxmax = g->zmaxmax;
  int rc;
  if (g->kind != MCB_GRID_VORONOI) {
    const int expect = g->n_rad * m.nj * g->n_az;
    if (expect != g->n_cells) return fail(h, MCB_ERR_BAD_ARG, "n_cells does not match n_rad*nz*n_az");
    h->gk = g->kind == MCB_GRID_CYL ? (g->l3D ? GK_CYL3D : GK_CYL2D) : (g->l3D ? GK_SPH3D : GK_SPH2D);
    if (!g->r_lim_2) return fail(h, MCB_ERR_BAD_ARG, "r_lim_2 missing");
    if ((rc = put(h, "r_lim_2", g->r_lim_2, (size_t)g->n_rad + 1, &m.r_lim_2))) return rc;
    if ((rc = put(h, "r_lim_3", g->r_lim_3, (size_t)g->n_rad + 1, &m.r_lim_3))) return rc;
    if ((rc = put(h, "r_lim", g->r_lim, (size_t)g->n_rad + 1, &m.r_lim))) return rc;      // distance_to_closest_wall_* only
    if (g->kind == MCB_GRID_CYL) {
      if (!g->z_lim || !g->zmax) return fail(h, MCB_ERR_BAD_ARG, "z_lim / zmax missing");
      if ((rc = put(h, "z_lim", g->z_lim, (size_t)g->n_rad * (g->nz + 2), &m.z_lim))) return rc;
      if ((rc = put(h, "zmax", g->zmax, (size_t)g->n_rad, &m.zmax))) return rc;
      // default vertical grid: z_lim(i,j) = (j-1)*cell_height(i), z_lim(i,nz+1) = zmax(i) (cylindrical_grid.f90:458-465).
      // If that holds bit-for-bit only cell_height(i) = z_lim(i,2) is needed on the device.
      bool regular = g->nz >= 2;
      for (int i = 0; i < g->n_rad && regular; ++i) {
        const double ch = g->z_lim[i + (size_t)g->n_rad * 1];
        for (int j = 1; j <= g->nz; ++j) if (g->z_lim[i + (size_t)g->n_rad * (j - 1)] != (double)(j - 1) * ch) { regular = false; break; }
        if (g->z_lim[i + (size_t)g->n_rad * g->nz] != g->zmax[i]) regular = false;
      }
      m.z_regular = regular ? 1 : 0;
      if (regular) {
        std::vector<double> ch((size_t)g->n_rad);
        for (int i = 0; i < g->n_rad; ++i) ch[i] = g->z_lim[i + (size_t)g->n_rad * 1];
        if ((rc = put(h, "cell_height", ch.data(), ch.size(), &m.cell_height))) return rc;
        CK(cudaStreamSynchronize(h->stream));
      }
    } else {
      if (!g->tan_theta_lim || !g->theta_lim) return fail(h, MCB_ERR_BAD_ARG, "tan_theta_lim / theta_lim missing");
      if ((rc = put(h, "tan_theta_lim", g->tan_theta_lim, (size_t)g->nz + 1, &m.tan_theta_lim))) return rc;
      if ((rc = put(h, "theta_lim", g->theta_lim, (size_t)g->nz + 1, &m.theta_lim))) return rc;
      if ((rc = put(h, "w_lim", g->w_lim, (size_t)g->nz + 1, &m.w_lim))) return rc;
      std::vector<double> cth((size_t)g->nz + 1);
      for (int j = 0; j <= g->nz; ++j) cth[j] = cos(g->theta_lim[j]);       // host libm, like cos_tab
      if ((rc = put(h, "cos_theta_lim", cth.data(), cth.size(), &m.cos_theta_lim))) return rc;
      CK(cudaStreamSynchronize(h->stream));
    }
    if (g->l3D) {
      if (!g->tan_phi_lim) return fail(h, MCB_ERR_BAD_ARG, "tan_phi_lim missing");
      if ((rc = put(h, "tan_phi_lim", g->tan_phi_lim, (size_t)g->n_az, &m.tan_phi_lim))) return rc;
      if ((rc = put(h, "sin_phi_lim", g->sin_phi_lim, (size_t)g->n_az, &m.sin_phi_lim))) return rc;
      if ((rc = put(h, "cos_phi_lim", g->cos_phi_lim, (size_t)g->n_az, &m.cos_phi_lim))) return rc;
    }
    // verify the closed-form numbering against the caller's cell maps
    if (g->cell_map_i && g->cell_map_j && g->cell_map_k && g->n_cells_tot > 0) {
      for (int id = 1; id <= g->n_cells_tot; ++id) {
        Cell c = cell_from_id(m, id);
        if (c.ri != g->cell_map_i[id - 1] || c.zj != g->cell_map_j[id - 1] || c.k != g->cell_map_k[id - 1] || cell_id(m, c) != id) {
          snprintf(h->err, sizeof h->err, "cell numbering mismatch at id %d: lib (%d,%d,%d) caller (%d,%d,%d)", id, c.ri, c.zj, c.k,
                   g->cell_map_i[id - 1], g->cell_map_j[id - 1], g->cell_map_k[id - 1]);
          return MCB_ERR_CELL_MAP;
        }
      }
    }
  } else {
    h->gk = GK_VOR;
    if (!g->vor_xyz || !g->vor_first || !g->vor_last || !g->neighbours_list) return fail(h, MCB_ERR_BAD_ARG, "Voronoi arrays missing");
    if ((rc = put(h, "vor_xyz", g->vor_xyz, (size_t)3 * g->n_cells, &m.vor_xyz))) return rc;
    if ((rc = put(h, "vor_h", g->vor_h, (size_t)g->n_cells, &m.vor_h))) return rc;
    if ((rc = put(h, "vor_first", g->vor_first, (size_t)g->n_cells, &m.vor_first))) return rc;
    if ((rc = put(h, "vor_last", g->vor_last, (size_t)g->n_cells, &m.vor_last))) return rc;
    if ((rc = put(h, "neigh", g->neighbours_list, (size_t)g->n_neighbours_tot, &m.neigh))) return rc;
    std::vector<float4> x32(g->n_cells);
    std::vector<uint8_t> fl(g->n_cells);
    for (int i = 0; i < g->n_cells; ++i) {
      x32[i] = make_float4((float)g->vor_xyz[3 * (size_t)i], (float)g->vor_xyz[3 * (size_t)i + 1], (float)g->vor_xyz[3 * (size_t)i + 2], 0.f);
      fl[i] = (uint8_t)((g->vor_was_cut && g->vor_was_cut[i] ? 1 : 0) | (g->vor_is_star && g->vor_is_star[i] ? 2 : 0) |
                        (g->vor_is_star_neighbour && g->vor_is_star_neighbour[i] ? 4 : 0));
    }
    if ((rc = put(h, "vor_xyz32", x32.data(), x32.size(), &m.vor_xyz32))) return rc;
    if ((rc = put(h, "vor_flags", fl.data(), fl.size(), &m.vor_flags))) return rc;
    // uniform grid over the seeds for point location (GeomVor::index; the reference's kd-tree, Voronoi.f90:1625-1645):
    // ~4 seeds per grid cell, ids ascending inside a cell (counting sort keeps the order of the ids)
    m.vg_start = nullptr; m.vg_items = nullptr;
    std::vector<int> vg_start, vg_items;
    if (g->n_cells >= 64) {
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      for (int i = 0; i < g->n_cells; ++i) for (int a = 0; a < 3; ++a) { const double q = g->vor_xyz[3 * (size_t)i + a]; lo[a] = std::min(lo[a], q); hi[a] = std::max(hi[a], q); }
      const double per_axis = std::cbrt((double)g->n_cells / 4.0);
      size_t ng = 1;
      for (int a = 0; a < 3; ++a) {
        const double ext = std::max(hi[a] - lo[a], 1e-300);
        m.vg_n[a] = std::max(1, std::min(1024, (int)per_axis));
        m.vg_lo[a] = lo[a]; m.vg_step[a] = ext / m.vg_n[a] * (1.0 + 1e-12); m.vg_inv[a] = 1.0 / m.vg_step[a];
        ng *= (size_t)m.vg_n[a];
      }
      auto cell_of = [&](int i) {
        size_t gidx = 0, mul = 1;
        for (int a = 0; a < 3; ++a) {
          const double q = std::floor((g->vor_xyz[3 * (size_t)i + a] - m.vg_lo[a]) * m.vg_inv[a]);
          const int c = q < 0.0 ? 0 : (q >= (double)m.vg_n[a] ? m.vg_n[a] - 1 : (int)q);
          gidx += mul * (size_t)c; mul *= (size_t)m.vg_n[a];
        }
        return gidx;
      };
      vg_start.assign(ng + 1, 0);
      for (int i = 0; i < g->n_cells; ++i) ++vg_start[cell_of(i) + 1];
      for (size_t c = 0; c < ng; ++c) vg_start[c + 1] += vg_start[c];
      vg_items.resize((size_t)g->n_cells);
      std::vector<int> fill(vg_start.begin(), vg_start.end() - 1);
      for (int i = 0; i < g->n_cells; ++i) vg_items[(size_t)fill[cell_of(i)]++] = i + 1;
      if ((rc = put(h, "vg_start", vg_start.data(), vg_start.size(), &m.vg_start))) return rc;
      if ((rc = put(h, "vg_items", vg_items.data(), vg_items.size(), &m.vg_items))) return rc;
    }
    CK(cudaStreamSynchronize(h->stream));       // x32 / fl are stack-owned
    memcpy(m.wall, g->wall_x, sizeof m.wall);
    m.cut_o_h = g->cutting_distance_o_h;
  }
  if ((rc = put(h, "volume", g->volume, (size_t)g->n_cells, &m.volume))) return rc;
  m.n_stars = g->n_stars;
  for (int i = 0; i < g->n_stars; ++i) {
    for (int a = 0; a < 4; ++a) m.star[i][a] = g->star_xyzr[4 * i + a];
    m.star_icell[i] = g->star_icell ? g->star_icell[i] : 0;
    m.star_out[i] = g->star_out_model ? g->star_out_model[i] : 0;
  }
  // default: no dark zone
  std::vector<uint8_t> dz((size_t)g->n_cells, 0);
  if ((rc = put(h, "dark", dz.data(), dz.size(), &m.dark))) return rc;
  h->host_dark = dz; h->kf_dark_stale = true;
  CK(cudaStreamSynchronize(h->stream));
  h->has_grid = true;
  return MCB_OK;
}

int mcfost_b200_upload_dark_zone(mcb_handle* h, const int32_t* l_dark_zone) {
  if (!h) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "upload_dark_zone before upload_grid");
  CK(cudaSetDevice(h->device));
  std::vector<uint8_t> dz((size_t)h->m.n_cells, 0);
  if (l_dark_zone) for (int i = 0; i < h->m.n_cells; ++i) dz[i] = l_dark_zone[i] != 0;
  h->host_dark = dz;
  h->kf_dark_stale = true;
  int rc;
  if ((rc = put(h, "dark", dz.data(), dz.size(), &h->m.dark))) return rc;
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_upload_opacity(mcb_handle* h, const mcb_opacity* o) {
  if (!h || !o) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "upload_opacity before upload_grid");
  CK(cudaSetDevice(h->device));
  DevModel& m = h->m;
  if (o->p_n_cells != 1 && o->p_n_cells != m.n_cells) return fail(h, MCB_ERR_BAD_ARG, "p_n_cells must be 1 or n_cells");
  if (o->p_n_lambda_pos != 1 && o->p_n_lambda_pos != o->n_lambda) return fail(h, MCB_ERR_BAD_ARG, "p_n_lambda_pos must be 1 or n_lambda");
  if (o->n_lambda < 1 || o->n_lambda > 8191) return fail(h, MCB_ERR_UNSUPPORTED, "n_lambda must be 1..8191 (13 bits of the packed packet state)");
  if (o->n_T < 2) return fail(h, MCB_ERR_BAD_ARG, "n_T < 2");
  h->mrw_ready = false;
  m.n_lambda = o->n_lambda; m.p_n_cells = o->p_n_cells; m.p_n_lambda_pos = o->p_n_lambda_pos; m.n_T = o->n_T;
  m.T_min = o->T_min;
  const size_t npl = (size_t)o->p_n_cells * o->n_lambda;
  const size_t npos = (size_t)(NANG + 1) * o->p_n_cells * o->p_n_lambda_pos;
  int rc;
  if (!o->kappa || !o->kappa_abs_LTE || !o->kappa_factor || !o->tab_albedo_pos) return fail(h, MCB_ERR_BAD_ARG, "opacity tables missing");
  if ((rc = put(h, "kappa", o->kappa, npl, &m.kappa))) return rc;
  if ((rc = put(h, "kappa_abs", o->kappa_abs_LTE, npl, &m.kappa_abs))) return rc;
  if ((rc = put(h, "kappa_factor", o->kappa_factor, (size_t)m.n_cells, &m.kappa_factor))) return rc;
  h->host_kappa_factor.assign(o->kappa_factor, o->kappa_factor + m.n_cells);
  h->kf_dark_stale = true;
  if ((rc = put(h, "albedo", o->tab_albedo_pos, npl, &m.albedo))) return rc;
  if ((rc = put(h, "gfac", o->tab_g_pos, npl, &m.gfac))) return rc;
  if ((rc = put(h, "prob_s11", o->prob_s11_pos, npos, &m.prob_s11))) return rc;
  if ((rc = put(h, "s11", o->tab_s11_pos, npos, &m.s11))) return rc;
  if ((rc = put(h, "s12", o->tab_s12_o_s11_pos, npos, &m.s12))) return rc;
  if ((rc = put(h, "s22", o->tab_s22_o_s11_pos, npos, &m.s22))) return rc;
  if ((rc = put(h, "s33", o->tab_s33_o_s11_pos, npos, &m.s33))) return rc;
  if ((rc = put(h, "s34", o->tab_s34_o_s11_pos, npos, &m.s34))) return rc;
  if ((rc = put(h, "s44", o->tab_s44_o_s11_pos, npos, &m.s44))) return rc;
  if ((rc = put(h, "logQ", o->log_Qcool_minus_extra_heating, (size_t)o->n_T * o->p_n_cells, &m.logQ))) return rc;
  if ((rc = put(h, "tab_Temp", o->tab_Temp, (size_t)o->n_T, &m.tab_Temp))) return rc;
  if ((rc = put(h, "kdB", o->kdB_dT_CDF, (size_t)o->n_lambda * o->n_T * o->p_n_cells, &m.kdB))) return rc;
  // cos(k*pi/nang_scatt), k = 0..180, with the host libm (scattering.f90:1470-1471 evaluates
  // cos((real(k,dp)-1)*pi/real(nang_scatt,dp)) per event; tabulated once here)
  std::vector<double> ct(NANG + 1);
  for (int k = 0; k <= NANG; ++k) ct[k] = cos(((double)k) * MCB_PI / (double)NANG);
  if ((rc = put(h, "cos_tab", ct.data(), ct.size(), &m.cos_tab))) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->has_op = true;
  return MCB_OK;
}

int mcfost_b200_upload_emission(mcb_handle* h, const mcb_emission* e) {
  if (!h || !e) return MCB_ERR_BAD_ARG;
  if (!h->has_op) return fail(h, MCB_ERR_STATE, "upload_emission before upload_opacity");
  CK(cudaSetDevice(h->device));
  DevModel& m = h->m;
  int rc;
  // frac_E_stars, frac_E_disk and prob_E_cell may be NULL when mcfost_b200_repartition_energie built them on the device
  const bool dev_tables = h->em_on_device && !e->frac_E_stars && !e->frac_E_disk && !e->prob_E_cell;
  if (!e->CDF_E_star || (!dev_tables && (!e->frac_E_stars || !e->frac_E_disk || !e->prob_E_cell))) return fail(h, MCB_ERR_BAD_ARG, "emission tables missing");
  if ((rc = put(h, "spec_cumul", e->spectre_emission_cumul, (size_t)m.n_lambda + 1, &m.spec_cumul))) return rc;
  if (!dev_tables) {
    if ((rc = put(h, "frac_star", e->frac_E_stars, (size_t)m.n_lambda, &m.frac_star))) return rc;
    if ((rc = put(h, "frac_disk", e->frac_E_disk, (size_t)m.n_lambda, &m.frac_disk))) return rc;
    if ((rc = put(h, "prob_E_cell", e->prob_E_cell, (size_t)(m.n_cells + 1) * m.n_lambda, &m.prob_E_cell))) return rc;
    h->em_on_device = false;
  }
  if ((rc = put(h, "CDF_E_star", e->CDF_E_star, (size_t)m.n_lambda * (m.n_stars + 1), &m.CDF_E_star))) return rc;
  if ((rc = put(h, "correct_E", e->correct_E_emission, (size_t)m.n_cells, &m.correct_E))) return rc;
  m.L_packet_th = e->L_packet_th; m.E_paquet = e->E_paquet; m.R_ISM = e->R_ISM;
  for (int a = 0; a < 3; ++a) m.cISM[a] = e->centre_ISM[a];
  CK(cudaStreamSynchronize(h->stream));
  h->has_em = true;
  return MCB_OK;
}

int mcfost_b200_upload_grains(mcb_handle* h, const mcb_grains* g) {
  if (!h || !g) return MCB_ERR_BAD_ARG;
  if (!h->has_op) return fail(h, MCB_ERR_STATE, "upload_grains before upload_opacity");
  CK(cudaSetDevice(h->device));
  DevModel& m = h->m;
  DevGrains& d = m.gr;
  memset(&d, 0, sizeof d);
  if (g->n_grains_tot < 1 || g->n_dens < 1 || !g->n_grains || !g->dust_density_o_n_grains) return fail(h, MCB_ERR_BAD_ARG, "grain tables missing");
  if (m.p_n_cells != 1 && g->n_dens != g->n_grains_tot) return fail(h, MCB_ERR_BAD_ARG, "lvariable_dust needs dust_density_o_n_grains(n_grains_tot, n_cells)");
  d.n_grains_tot = g->n_grains_tot; d.n_dens = g->n_dens;
  d.LTE_s = g->grain_RE_LTE_start; d.LTE_e = g->grain_RE_LTE_end;
  d.nLTE_s = g->grain_RE_nLTE_start; d.nLTE_e = g->grain_RE_nLTE_end;
  d.nRE_s = g->grain_nRE_start; d.nRE_e = g->grain_nRE_end;
  const size_t K = g->n_grains_tot, nl = m.n_lambda, nc = m.n_cells, nT = m.n_T;
  const size_t k1 = d.nLTE_e >= d.nLTE_s && d.nLTE_s > 0 ? (size_t)(d.nLTE_e - d.nLTE_s + 1) : 0;
  const size_t k2 = d.nRE_e >= d.nRE_s && d.nRE_s > 0 ? (size_t)(d.nRE_e - d.nRE_s + 1) : 0;
  int rc;
  if ((rc = put(h, "g_zone", g->grain_zone, K, &d.zone))) return rc;
  if ((rc = put(h, "g_n_grains", g->n_grains, K, &d.n_grains))) return rc;
  if ((rc = put(h, "g_dd", g->dust_density_o_n_grains, (size_t)g->n_dens * nc, &d.dd))) return rc;
  if ((rc = put(h, "g_C_abs", g->C_abs, K * nl, &d.C_abs))) return rc;
  if ((rc = put(h, "g_C_abs_norm", g->C_abs_norm, K * nl, &d.C_abs_norm))) return rc;
  if ((rc = put(h, "g_C_sca", g->C_sca, K * nl, &d.C_sca))) return rc;
  if ((rc = put(h, "g_tab_g", g->tab_g, K * nl, &d.tab_g))) return rc;
  const size_t ns = (size_t)(NANG + 1) * K * nl;
  if ((rc = put(h, "g_prob_s11", g->prob_s11, ns, &d.prob_s11))) return rc;
  if ((rc = put(h, "g_s11", g->tab_s11, ns, &d.s11))) return rc;
  if ((rc = put(h, "g_s12", g->tab_s12, ns, &d.s12))) return rc;
  if ((rc = put(h, "g_s22", g->tab_s22, ns, &d.s22))) return rc;
  if ((rc = put(h, "g_s33", g->tab_s33, ns, &d.s33))) return rc;
  if ((rc = put(h, "g_s34", g->tab_s34, ns, &d.s34))) return rc;
  if ((rc = put(h, "g_s44", g->tab_s44, ns, &d.s44))) return rc;
  if ((rc = put(h, "g_ksca_CDF", g->ksca_CDF, (K + 1) * (size_t)m.p_n_cells * nl, &d.ksca_CDF))) return rc;
  if ((rc = put(h, "g_kappa_abs_nLTE", g->kappa_abs_nLTE, (size_t)m.p_n_cells * nl, &d.kappa_abs_nLTE))) return rc;
  if ((rc = put(h, "g_kabs_nLTE_CDF", g->kabs_nLTE_CDF, (k1 + 1) * nc * nl, &d.kabs_nLTE_CDF))) return rc;
  if ((rc = put(h, "g_logE", g->log_E_em_1grain, k1 * nT, &d.logE))) return rc;
  if ((rc = put(h, "g_kdB", g->kdB_dT_1grain_nLTE_CDF, nl * k1 * nT, &d.kdB))) return rc;
  if ((rc = put(h, "g_kappa_abs_RE", g->kappa_abs_RE, nc * nl, &d.kappa_abs_RE))) return rc;
  if ((rc = put(h, "g_proba_abs_RE", g->proba_abs_RE, nc * nl, &d.proba_abs_RE))) return rc;
  if ((rc = put(h, "g_P_LTE", g->Proba_abs_RE_LTE, nc * nl, &d.P_LTE))) return rc;
  if ((rc = put(h, "g_P_LTE_p_nLTE", g->Proba_abs_RE_LTE_p_nLTE, nc * nl, &d.P_LTE_p_nLTE))) return rc;
  if ((rc = put(h, "g_logE_nRE", g->log_E_em_1grain_nRE, k2 * nT, &d.logE_nRE))) return rc;
  if ((rc = put(h, "g_kdB_nRE", g->kdB_dT_1grain_nRE_CDF, nl * k2 * nT, &d.kdB_nRE))) return rc;
  if ((rc = put(h, "g_l_RE", g->l_RE, k2 * nc, &d.l_RE))) return rc;
  if ((rc = put(h, "g_J0", g->J0, nc * nl, &d.J0))) return rc;
  {
    const size_t k0 = d.LTE_e >= d.LTE_s && d.LTE_s > 0 ? (size_t)(d.LTE_e - d.LTE_s + 1) : 0;
    if ((rc = put(h, "g_kdB_LTE", g->kdB_dT_1grain_LTE_CDF, nl * k0 * nT, &d.kdB_LTE))) return rc;
  }
  CK(cudaStreamSynchronize(h->stream));
  h->gr_host = *g;
  h->has_gr = true;
  return MCB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Shared-memory staging plan (SmemLayout): enabled when the dust is not cell-dependent and
// two blocks' worth of tables fit next to L1 on one SM.
static void compute_smem_layout(mcb_handle* h, int /*p_lambda_in*/) {
  DevModel& m = h->m;
  SmemLayout L; memset(&L, 0, sizeof L);
  if (m.p_n_cells == 1 && h->gk != GK_VOR) {
    int off = 0;
    auto take = [&](int n_words) { int o = off; off += n_words; return o; };
    L.r_lim_2 = take(m.n_rad + 1);
    if (m.kind == MCB_GRID_CYL) { L.zmax = take(m.n_rad); L.zl = take(m.z_regular ? m.n_rad : m.n_rad * (m.nz + 2)); }
    else L.tan_theta = take(m.nz + 1);
    if (m.l3D) L.tan_phi = take(m.n_az);
    L.kappa = take(m.n_lambda); L.kappa_abs = take(m.n_lambda);
    L.albedo = take((m.n_lambda + 1) / 2); L.gfac = take((m.n_lambda + 1) / 2);
    L.logQ = take(m.n_T);
    // kdB_dT_CDF (40 KB for ref4.1) is NOT staged: measured +3.7 % with it in global memory, because the
    // shared memory it would take is worth more as L1 for kappa_factor / volume / tally lines
    L.kdB = -1;
    L.cos_tab = take(NANG + 1); L.prob_s11 = take((NANG + 2) / 2);
    L.spec_cumul = take(m.n_lambda + 1); L.frac_star = take(m.n_lambda); L.frac_disk = take(m.n_lambda);
    L.total_words = off;
    L.enabled = (off * 8 <= 96 * 1024) && m.logQ && m.kdB && m.spec_cumul;
  }
  m.sm = L;
}


// ===========================================================================
// Modified random walk: tables (MRW.f90:16-54 zeta; mean opacities, the job of compute_Planck_opacities
// diffusion.f90:631-693, see DESIGN.md "MRW")
// ===========================================================================
// One thread per (temperature index, p_icell): A = Sum p/(kappa(1-a)), B = Sum p/(kappa(1-a) kappa(1-a g)),
// C = Sum p kappa_abs_LTE/(kappa(1-a)) with p = increments of kdB_dT_CDF(:, T, p_icell).
__global__ void mrw_means_kernel(const __grid_constant__ DevModel m, double* A, double* B, double* Cc) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= (int64_t)m.n_T * m.p_n_cells) return;
  const int t = (int)(q % m.n_T), pc = (int)(q / m.n_T);
  const double* cdf = m.kdB + (size_t)m.n_lambda * ((size_t)t + (size_t)m.n_T * pc);
  double a_ = 0.0, b_ = 0.0, c_ = 0.0, prev = 0.0;
  for (int l = 0; l < m.n_lambda; ++l) {
    const double cur = cdf[l];
    const double pl = cur - prev;
    prev = cur;
    const size_t pl_i = (size_t)pc + (size_t)m.p_n_cells * l;
    const double kap = m.kappa[pl_i], alb = (double)m.albedo[pl_i];
    const double gg = m.gfac ? (double)m.gfac[pl_i] : 0.0;
    const double k_abs = kap * (1.0 - alb), k_tr = kap * (1.0 - alb * gg);
    if (!(pl > 0.0) || !(k_abs > 0.0)) continue;
    a_ = a_ + pl / k_abs;
    b_ = b_ + pl / (k_abs * k_tr);
    c_ = c_ + pl * m.kappa_abs[pl_i] / k_abs;
  }
  A[q] = a_; B[q] = b_; Cc[q] = c_;
}

static int ensure_mrw_tables(mcb_handle* h) {
  if (h->mrw_ready) return MCB_OK;
  DevModel& m = h->m;
  if (!m.kdB || !m.kappa || !m.kappa_abs || !m.albedo) return fail(h, MCB_ERR_BAD_ARG, "lMRW: thermal / opacity tables missing");
  // zeta(y) = 2 Sum_{n>=1} (-1)^(n+1) y^(n^2), y = (i-1)/(n-1) (MRW.f90:16-54), made monotone by a running maximum
  // (beyond y ~ 0.95 the series is 1 to rounding)
  std::vector<double> zeta((size_t)N_ZETA, 0.0);
  for (int i = 1; i <= N_ZETA; ++i) {
    const double y = (double)(i - 1) / (double)(N_ZETA - 1);
    double z = 0.0;
    if (i == N_ZETA) z = 0.5;
    else {
      int j = 0;
      for (;;) {
        j = j + 1;
        const double term = pow(y, (double)j * (double)j);
        if (term == 0.0) break;
        if (j % 2 == 0) z = z - term; else z = z + term;
      }
    }
    zeta[i - 1] = z * 2.0;
  }
  for (int i = 1; i < N_ZETA; ++i) if (zeta[i] < zeta[i - 1]) zeta[i] = zeta[i - 1];
  int rc;
  if ((rc = put(h, "zeta", zeta.data(), zeta.size(), &m.zeta))) return rc;
  const size_t n = (size_t)m.n_T * m.p_n_cells;
  double *A = nullptr, *B = nullptr, *Cc = nullptr;
  if ((rc = reserve(h, "mrw_A", n, &A)) || (rc = reserve(h, "mrw_B", n, &B)) || (rc = reserve(h, "mrw_C", n, &Cc))) return rc;
  mrw_means_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(m, A, B, Cc);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));          // zeta is a host vector
  m.mrw_A = A; m.mrw_B = B; m.mrw_C = Cc;
  h->mrw_ready = true;
  return MCB_OK;
}

extern "C" int mcfost_b200_mrw_tables(mcb_handle* h, double* A, double* B, double* Cc) {
  if (!h) return MCB_ERR_BAD_ARG;
  if (!h->has_op) return fail(h, MCB_ERR_STATE, "mrw_tables before upload_opacity");
  CK(cudaSetDevice(h->device));
  int rc = ensure_mrw_tables(h);
  if (rc) return rc;
  const size_t bytes = (size_t)h->m.n_T * h->m.p_n_cells * sizeof(double);
  if (A) CK(cudaMemcpyAsync(A, h->m.mrw_A, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (B) CK(cudaMemcpyAsync(B, h->m.mrw_B, bytes, cudaMemcpyDeviceToHost, h->stream));
  if (Cc) CK(cudaMemcpyAsync(Cc, h->m.mrw_C, bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

static int setup_tallies(mcb_handle* h, const mcb_run_params* r, bool lxJ, bool rt1, int n_type_flux, bool rt2) {
  DevModel& m = h->m;
  const int n_sed = m.n_lambda * r->N_thet * r->N_phi;
  bool realloc_ = (h->n_tally == 0) || (h->lay_xJ != lxJ) || (h->lay_nsed != n_sed);
  TallyLayout L;
  L.xKJ = 0;
  L.xJ = m.n_cells;
  L.n_env = L.xJ + (lxJ ? (int64_t)m.n_cells * m.n_lambda : 0);
  L.sed = L.n_env + m.n_lambda;
  L.n_sed = n_sed;
  L.stats = L.sed + 9 * (int64_t)n_sed;
  L.E_abs_nRE = L.stats + 12;
  L.total = L.E_abs_nRE + 1;
  m.lay = L;
  int rc;
  if ((rc = reserve(h, "tally", (size_t)L.total, &m.tally))) return rc;
  if ((rc = reserve(h, "xT_ech", (size_t)m.n_cells, &m.xT_ech))) return rc;
  const int n_rt = r->RT_n_incl * r->RT_n_az;
  const int64_t n_xI = rt1 ? (int64_t)N_AZ_RT * 2 * n_type_flux * n_rt * m.n_cells : 0;
  if (n_xI != h->n_xI) realloc_ = true;
  if ((rc = reserve(h, "xI", (size_t)n_xI, &m.xI))) return rc;
  const int64_t n_Is = rt2 ? (int64_t)n_type_flux * r->n_theta_I * r->n_phi_I * m.n_cells : 0;
  if (n_Is != h->n_Ispec) realloc_ = true;
  if ((rc = reserve(h, "I_spec", (size_t)n_Is, &m.I_spec))) return rc;
  if ((rc = reserve(h, "I_spec_star", (size_t)(rt2 ? m.n_cells : 0), &m.I_spec_star))) return rc;
  h->n_Ispec = n_Is;
  // per-grain temperature indices (thermal_emission.f90:163-165,186-188)
  const int64_t n_1g = (h->has_gr && r->lRE_nLTE) ? (int64_t)(m.gr.nLTE_e - m.gr.nLTE_s + 1) * m.n_cells : 0;
  const int64_t n_1g_nRE = (h->has_gr && r->lnRE) ? (int64_t)(m.gr.nRE_e - m.gr.nRE_s + 1) * m.n_cells : 0;
  if (n_1g != h->n_1g || n_1g_nRE != h->n_1g_nRE) realloc_ = true;
  if ((rc = reserve(h, "xT_1g", (size_t)n_1g, &m.gr.xT_1g))) return rc;
  if ((rc = reserve(h, "xT_1g_nRE", (size_t)n_1g_nRE, &m.gr.xT_1g_nRE))) return rc;
  h->n_1g = n_1g; h->n_1g_nRE = n_1g_nRE;
  if ((rc = reserve(h, "work", (size_t)(48 + 2 * r->n_photons_loop), &m.work))) return rc;
  h->n_tally = L.total; h->n_xI = n_xI; h->lay_xJ = lxJ; h->lay_nsed = n_sed; h->n_type_flux = n_type_flux;
  h->rt1_n_rt = rt1 ? n_rt : 0; h->rt1_pola = r->lsepar_pola ? 1 : 0; h->rt1_contrib = r->lsepar_contrib ? 1 : 0;
  if (realloc_ || r->reset_tallies) {
    CK(cudaMemsetAsync(m.tally, 0, (size_t)L.total * sizeof(double), h->stream));
    if (n_xI) CK(cudaMemsetAsync(m.xI, 0, (size_t)n_xI * sizeof(float), h->stream));
    if (n_Is) { CK(cudaMemsetAsync(m.I_spec, 0, (size_t)n_Is * sizeof(float), h->stream)); CK(cudaMemsetAsync(m.I_spec_star, 0, (size_t)m.n_cells * sizeof(float), h->stream)); }
    fill_int_kernel<<<256, 256, 0, h->stream>>>(m.xT_ech, m.n_cells, 2);      // xT_ech = 2, thermal_emission.f90:119,2164
    if (n_1g) fill_int_kernel<<<256, 256, 0, h->stream>>>(m.gr.xT_1g, n_1g, 2);
    if (n_1g_nRE) fill_int_kernel<<<256, 256, 0, h->stream>>>(m.gr.xT_1g_nRE, n_1g_nRE, 2);
  }
  CK(cudaMemsetAsync(m.tally + L.E_abs_nRE, 0, sizeof(double), h->stream));      // E_abs_nRE = 0.0 at every call (dust_transfer.f90:505)
  CK(cudaMemsetAsync(m.work, 0, (size_t)(48 + 2 * r->n_photons_loop) * sizeof(unsigned long long), h->stream));
  h->n_photons_loop_alloc = r->n_photons_loop;
  CK(cudaMemsetAsync(m.work + 1, 0xFF, sizeof(unsigned long long), h->stream));      // work[1] = ~0: "counter not dry yet"
  return MCB_OK;
}

extern "C" {

int mcfost_b200_launch(mcb_handle* h, const mcb_run_params* r) {
  if (!h || !r) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op || !h->has_em) return fail(h, MCB_ERR_STATE, "run before upload_grid/opacity/emission");
  CK(cudaSetDevice(h->device));
  DevModel& m = h->m;
  // ---- modes this library implements; everything else fails loudly ----
  {
    const mcb_grains& g = h->gr_host;
    const bool need_gr = r->lscattering_method1 || (!r->lonly_LTE && !r->lmono) || (r->low_mem_th_emission && !r->lmono);
    if (need_gr && !h->has_gr) return fail(h, MCB_ERR_STATE, "per-grain mode (method 1 / nLTE / nRE) before upload_grains");
    if (r->lscattering_method1) {
      if (!g.C_sca) return fail(h, MCB_ERR_BAD_ARG, "method 1: C_sca missing");
      if (r->low_mem_scattering ? (m.p_n_cells == 1 && !g.grain_zone) : !g.ksca_CDF) return fail(h, MCB_ERR_BAD_ARG, "method 1: grain_zone / ksca_CDF missing");
      if (r->lmethod_aniso1 ? !g.prob_s11 : !g.tab_g) return fail(h, MCB_ERR_BAD_ARG, "method 1: prob_s11 / tab_g missing");
      if (r->lsepar_pola && r->lmethod_aniso1 && (!g.tab_s11 || !g.tab_s12 || !g.tab_s22 || !g.tab_s33 || !g.tab_s34 || !g.tab_s44)) return fail(h, MCB_ERR_BAD_ARG, "method 1: per-grain Mueller tables missing");
      if ((!r->letape_th) && (r->lscatt_ray_tracing1 || r->lscatt_ray_tracing2)) return fail(h, MCB_ERR_UNSUPPORTED, "ray-tracing accumulators need scattering method 2 (dust_ray_tracing.f90)");
    }
    if (r->low_mem_th_emission && !r->lmono && (!g.kdB_dT_1grain_LTE_CDF || !g.C_abs || g.grain_RE_LTE_start < 1 || g.grain_RE_LTE_end < g.grain_RE_LTE_start))
      return fail(h, MCB_ERR_BAD_ARG, "low_mem_th_emission: kdB_dT_1grain_LTE_CDF / C_abs / LTE grain range missing");
    if (r->lweight_emission && !m.correct_E) return fail(h, MCB_ERR_BAD_ARG, "lweight_emission: correct_E_emission missing (upload_emission)");
    if (r->lspot && (!r->tab_lambda || !(r->star1_T > 0.0) || !(r->T_spot > 0.0f))) return fail(h, MCB_ERR_BAD_ARG, "lspot: tab_lambda / star1_T / T_spot missing");
    if (!r->lonly_LTE && !r->lmono) {
      const bool mixed = !r->lonly_nLTE;
      if (!g.C_abs_norm || !g.J0) return fail(h, MCB_ERR_BAD_ARG, "nLTE / nRE: C_abs_norm / J0 missing");
      if (!(r->letape_th ? r->lxJ_abs_step1 : r->lxJ_abs)) return fail(h, MCB_ERR_BAD_ARG, "nLTE / nRE re-emission reads xJ_abs: lxJ_abs(_step1) must be on");
      if (r->lRE_nLTE || r->lonly_nLTE) {
        if (g.grain_RE_nLTE_start < 1 || g.grain_RE_nLTE_end < g.grain_RE_nLTE_start || !g.log_E_em_1grain || !g.kdB_dT_1grain_nLTE_CDF) return fail(h, MCB_ERR_BAD_ARG, "nLTE tables missing");
        if (r->low_mem_th_emission_nLTE ? (!g.C_abs || !g.kappa_abs_nLTE) : !g.kabs_nLTE_CDF) return fail(h, MCB_ERR_BAD_ARG, "nLTE grain-selection tables missing");
        if (!r->lRE_nLTE) return fail(h, MCB_ERR_BAD_ARG, "lonly_nLTE without lRE_nLTE");
      }
      if (r->lnRE) {
        if (g.grain_nRE_start < 1 || g.grain_nRE_end < g.grain_nRE_start || !g.log_E_em_1grain_nRE || !g.kdB_dT_1grain_nRE_CDF || !g.l_RE || !g.kappa_abs_RE || !g.proba_abs_RE || !g.C_abs)
          return fail(h, MCB_ERR_BAD_ARG, "nRE tables missing");
        if (r->lRE_nLTE && !g.kappa_abs_nLTE) return fail(h, MCB_ERR_BAD_ARG, "kappa_abs_nLTE missing");
      }
      if (mixed && (!g.Proba_abs_RE_LTE || !g.Proba_abs_RE_LTE_p_nLTE)) return fail(h, MCB_ERR_BAD_ARG, "Proba_abs_RE_LTE(_p_nLTE) missing");
      if (mixed && !r->lnRE && !r->lRE_nLTE) return fail(h, MCB_ERR_BAD_ARG, "lonly_LTE = 0 without lRE_nLTE or lnRE");
    }
  }
  const bool mc_maps = r->lmono0 && r->loutput_mc;
  if (mc_maps && (r->npix_x < 1 || r->npix_y < 1 || !(r->map_size > 0.0))) return fail(h, MCB_ERR_BAD_ARG, "loutput_mc needs npix_x, npix_y, map_size");
  if (r->lorigine && (r->capt_interet < 1 || r->capt_interet > r->N_thet)) return fail(h, MCB_ERR_BAD_ARG, "capt_interet out of range");
  if (r->lscatt_ray_tracing2 && (m.l3D || h->gk == GK_VOR)) return fail(h, MCB_ERR_UNSUPPORTED, "rt2 is 2D only (radiation_field.f90:91)");
  if (r->lscatt_ray_tracing2 && (r->n_theta_I < 1 || r->n_phi_I < 1)) return fail(h, MCB_ERR_BAD_ARG, "n_theta_I / n_phi_I");
  if (r->n_photons_loop < 1 || r->nnfot1_start < 1 || r->n_photons2 < 0) return fail(h, MCB_ERR_BAD_ARG, "bad packet budget");
  if (r->lambda_in < 1 || r->lambda_in > m.n_lambda || r->p_lambda_in < 1 || r->p_lambda_in > m.p_n_lambda_pos) return fail(h, MCB_ERR_BAD_ARG, "lambda index out of range");
  if (r->n_ranks < 1 || r->rank < 0 || r->rank >= r->n_ranks) return fail(h, MCB_ERR_BAD_ARG, "bad rank / n_ranks");
  if (r->lmethod_aniso1 && (!m.prob_s11)) return fail(h, MCB_ERR_BAD_ARG, "prob_s11_pos missing for lmethod_aniso1");
  if (!r->lmethod_aniso1 && !m.gfac) return fail(h, MCB_ERR_BAD_ARG, "tab_g_pos missing for HG scattering");
  if (r->lsepar_pola && r->lmethod_aniso1 && (!m.s12 || !m.s22 || !m.s33 || !m.s34 || !m.s44)) return fail(h, MCB_ERR_BAD_ARG, "Mueller tables missing for lsepar_pola");
  if (r->letape_th && (!m.logQ || !m.kdB || !m.spec_cumul)) return fail(h, MCB_ERR_BAD_ARG, "thermal tables missing");
  if (r->N_thet < 1 || r->N_phi < 1 || r->capt_sup < 1 || r->capt_sup > r->N_thet) return fail(h, MCB_ERR_BAD_ARG, "N_thet, N_phi >= 1 and 1 <= capt_sup <= N_thet");
  // the per-cell Mueller tables are read with the packet's own wavelength (scattering.f90:1333-1341): with the
  // single-wavelength layout (p_n_lambda_pos = 1) that index only exists for lambda = 1
  if (r->lsepar_pola && r->lmethod_aniso1 && !r->lscattering_method1 && m.p_n_lambda_pos == 1 && m.n_lambda > 1 && !r->lmono)
    return fail(h, MCB_ERR_UNSUPPORTED, "lsepar_pola with p_n_lambda_pos = 1 needs a monochromatic call (Mueller tables hold one wavelength)");
  if (r->lMRW) {
    if (!r->letape_th || r->lmono || !r->lonly_LTE || r->low_mem_th_emission || r->lxJ_abs_step1 || r->lscattering_method1)
      return fail(h, MCB_ERR_UNSUPPORTED, "lMRW: thermal step with lonly_LTE, scattering method 2 and no xJ_abs only");
    if (h->gk != GK_VOR && !m.r_lim) return fail(h, MCB_ERR_BAD_ARG, "lMRW: r_lim missing");
    if ((h->gk == GK_SPH2D || h->gk == GK_SPH3D) && !m.w_lim) return fail(h, MCB_ERR_BAD_ARG, "lMRW: w_lim missing");
    if ((h->gk == GK_CYL3D || h->gk == GK_SPH3D) && (!m.sin_phi_lim || !m.cos_phi_lim)) return fail(h, MCB_ERR_BAD_ARG, "lMRW: sin_phi_lim / cos_phi_lim missing");
    const int rcm = ensure_mrw_tables(h);
    if (rcm) return rcm;
  }
  const bool rt1 = (!r->letape_th) && r->lscatt_ray_tracing1;
  const int n_rt = r->RT_n_incl * r->RT_n_az;
  if (rt1) {
    if (n_rt < 1 || n_rt > MAX_RT) return fail(h, MCB_ERR_UNSUPPORTED, "rt1 needs 1..16 observer directions");
    if (!r->tab_u_rt || !r->tab_v_rt || !r->tab_w_rt || !m.s11) return fail(h, MCB_ERR_BAD_ARG, "rt1 direction / s11 tables missing");
  }
  DevRun dr;
  memset(&dr, 0, sizeof dr);
  dr.lambda_in = r->lambda_in; dr.p_lambda_in = r->p_lambda_in; dr.n_photons2 = r->n_photons2;
  dr.nnfot1_start = r->nnfot1_start; dr.n_photons_loop = r->n_photons_loop; dr.n_phot_lim = r->n_phot_lim;
  dr.letape_th = r->letape_th; dr.lmono = r->lmono; dr.lsepar_pola = r->lsepar_pola; dr.lsepar_contrib = r->lsepar_contrib;
  dr.lmethod_aniso1 = r->lmethod_aniso1; dr.lisotropic = r->lisotropic;
  dr.l_sym_centrale = r->l_sym_centrale; dr.l_sym_axiale = r->l_sym_axiale;
  dr.rt1 = rt1;
  dr.rt2 = ((!r->letape_th) && !rt1 && r->lscatt_ray_tracing2) ? 1 : 0;
  dr.lmono0 = r->lmono0; dr.n_theta_I = r->n_theta_I; dr.n_phi_I = r->n_phi_I;
  dr.lscattering_method1 = r->lscattering_method1; dr.low_mem_scattering = r->low_mem_scattering;
  dr.lonly_LTE = (r->lonly_LTE || r->lmono) ? 1 : 0;      // lmono: no absorption event ever happens (forced scattering)
  dr.lonly_nLTE = r->lonly_nLTE; dr.lRE_nLTE = r->lRE_nLTE; dr.lnRE = r->lnRE; dr.low_mem_nLTE = r->low_mem_th_emission_nLTE;
  dr.lxJ = r->letape_th ? r->lxJ_abs_step1 : r->lxJ_abs;
  dr.N_thet = r->N_thet; dr.N_phi = r->N_phi; dr.capt_sup = r->capt_sup;
  dr.n_stokes = r->lsepar_pola ? 4 : 1;
  dr.n_type_flux = dr.n_stokes + (r->lsepar_contrib ? 4 : 0);      // init_mcfost.f90:1604-1616
  dr.n_rt = rt1 ? n_rt : 0; dr.RT_n_incl = r->RT_n_incl; dr.RT_n_az = r->RT_n_az;
  for (int i = 0; i < dr.n_rt; ++i) {
    dr.rt_u[i] = r->tab_u_rt[i]; dr.rt_v[i] = r->tab_v_rt[i]; dr.rt_w[i] = r->tab_w_rt[i % r->RT_n_incl];
  }
  dr.seed = r->seed; dr.call_index = r->call_index;
  dr.rank = r->rank; dr.n_ranks = r->n_ranks;
  // chunks nnfot1_start..n_photons_loop dealt round-robin: rank owns chunks with (c-1) % n_ranks == rank
  int n_local = 0;
  for (int c = r->nnfot1_start; c <= r->n_photons_loop; ++c) if (((c - 1) % r->n_ranks) == r->rank) ++n_local;
  dr.n_local_chunks = n_local;
  if (n_local > 32767 * 65536) return fail(h, MCB_ERR_UNSUPPORTED, "too many chunks");
  dr.count_sent = (r->letape_th || r->lmono0 || r->lcount_sent) ? 1 : 0;      // dust_transfer.f90:503-518
  const double lim = ceil((double)r->n_phot_lim);
  dr.sent_lim = (lim >= 1.8e19) ? ~0ull : (lim <= 0 ? 0ull : (unsigned long long)lim);
  dr.n_per_chunk = (unsigned long long)r->n_photons2 < dr.sent_lim ? (unsigned long long)r->n_photons2 : dr.sent_lim;
  dr.n_packets_total = dr.count_sent ? (unsigned long long)n_local * dr.n_per_chunk : 0ull;
  dr.nb_proc_equiv = (double)r->n_ranks;
  dr.low_mem_th = (r->low_mem_th_emission && !r->lmono) ? 1 : 0;
  dr.lweight_emission = r->lweight_emission; dr.lspot = r->lspot; dr.lxN = r->lxN_abs ? 1 : 0;
  if (r->lspot) {      // dust_transfer.f90:1101-1107, evaluated once (the reference recomputes it per packet)
    const double PI_ = MCB_PI;
    dr.z_spot = (float)cos((double)(r->theta_spot / 180.0f) * PI_);
    dr.x_spot = (float)(sin((double)(r->theta_spot / 180.0f) * PI_) * cos((double)(r->phi_spot / 180.0f) * PI_));
    dr.y_spot = (float)(sin((double)(r->theta_spot / 180.0f) * PI_) * sin((double)(r->phi_spot / 180.0f) * PI_));
    dr.cos_thet_spot = sqrtf(1.0f - r->surf_fraction_spot);
    dr.T_spot = r->T_spot; dr.star1_T = r->star1_T;
    int rc2;
    if ((rc2 = put(h, "tab_lambda", r->tab_lambda, (size_t)m.n_lambda, &m.tab_lambda))) return rc2;
  }
  dr.mc_maps = mc_maps ? 1 : 0; dr.lorigine = r->lorigine; dr.capt_interet = r->capt_interet;
  dr.lonly_capt_interet = r->lonly_capt_interet; dr.capt_inf = r->capt_inf;
  dr.npix_x = r->npix_x; dr.npix_y = r->npix_y; dr.l_sym_ima = r->l_sym_ima;
  dr.zoom = (double)r->zoom; dr.map_size = r->map_size; dr.cos_disk = r->cos_disk; dr.sin_disk = r->sin_disk;
  dr.capt_full = (mc_maps || r->lorigine || r->lonly_capt_interet) ? 1 : 0;
  // (the flight-start slab of capteur_full is not part of a parked packet: no hand-over in those modes)
  dr.park_enable = 0;      // set by mcb_launch_mc for the kernels that hand their last packets over
  dr.patience = 8;
  dr.park_live = 128;     // measured (profiles/r02_latency_and_tail.md section 7): hand-over at 96 .. 192 live packets per SM is a flat optimum with the round's final packet-per-warp kernel (48 before its instruction diet)
  dr.debug_abort_dry = 0;
  dr.patience_dry = 1; dr.drain_live_dry = 96;
  dr.min_chunk = 32;
#ifdef MCB_DEV
  { const char* e = getenv("MCB_MIN_CHUNK"); if (e && atoi(e) > 0 && atoi(e) <= 32) dr.min_chunk = atoi(e); }
  { const char* e = getenv("MCB_PATIENCE_DRY"); if (e && atoi(e) >= 0) dr.patience_dry = atoi(e); }
  { const char* e = getenv("MCB_DRAIN_LIVE_DRY"); if (e && atoi(e) > 0) dr.drain_live_dry = atoi(e); }      // development builds only: a science library does not change its results on an environment variable
  { const char* e = getenv("MCB_PATIENCE"); if (e && atoi(e) > 0 && atoi(e) <= 4096) dr.patience = atoi(e); }
  { const char* e = getenv("MCB_PARK_LIVE"); if (e && atoi(e) > 0 && atoi(e) <= 256) dr.park_live = atoi(e); }
  { const char* e = getenv("MCB_DEBUG_ABORT_DRY"); dr.debug_abort_dry = (e && e[0] == '1') ? 1 : 0; }   // profiling aid: tallies are incomplete
#endif
  dr.lism = r->lISM_loop ? 1 : 0;
  if (dr.lism) {
    if (r->letape_th) return fail(h, MCB_ERR_BAD_ARG, "lISM_loop is a side loop of the SED step, not of the thermal step");
    if (!(m.R_ISM > 0.0)) return fail(h, MCB_ERR_BAD_ARG, "lISM_loop: R_ISM missing (upload_emission)");
    dr.rt1 = 0; dr.rt2 = 0; dr.n_rt = 0;      // lscatt_ray_tracing1/2 are switched off around the loop (:942-945)
    dr.count_sent = 0; dr.n_packets_total = 0ull;
  }
  dr.lMRW = r->lMRW ? 1 : 0;
  dr.gamma_MRW = (r->gamma_MRW > 0.0f) ? (double)r->gamma_MRW : 2.0;
  // concurrency window: immediate re-emission reads RUNNING tallies, so the packets in flight are kept a small
  // fraction of the packets already sent (a host run keeps nb_proc in flight); floor = one warp per block
  dr.inflight_floor = 32u;
  dr.inflight_frac_per_block = ((r->max_inflight_fraction > 0.0f) ? fminf(r->max_inflight_fraction, 1.0f) : 1.0f / 16.0f);      // the fraction itself; mc_kernel.cu divides it by its block count
  int rc = setup_tallies(h, r, dr.lxJ != 0, rt1, dr.n_type_flux, dr.rt2 != 0);
  if (rc) return rc;
  {   // capteur extras: photon maps of this wavelength, origin tallies, flight-start slab
    const int64_t n_map = mc_maps ? (int64_t)r->npix_x * r->npix_y * r->N_thet * r->N_phi * dr.n_type_flux : 0;
    const int64_t n_org = r->lorigine ? (int64_t)m.n_lambda * (m.n_cells + 1) : 0;
    const bool fresh = (n_map != h->n_map) || (n_org != h->n_org) || r->reset_tallies;
    if ((rc = reserve(h, "smap", (size_t)n_map, &m.smap))) return rc;
    if ((rc = reserve(h, "origin", (size_t)n_org, &m.star_origin))) return rc;
    m.disk_origin = m.star_origin + m.n_lambda;
    const int64_t n_xN = r->lxN_abs ? (int64_t)m.n_cells * (r->letape_th ? 1 : m.n_lambda) : 0;
    const bool fresh_xN = (n_xN != h->n_xN) || r->reset_tallies;
    if ((rc = reserve(h, "xN", (size_t)n_xN, &m.xN))) return rc;
    h->n_xN = n_xN;
    if (fresh_xN && n_xN) CK(cudaMemsetAsync(m.xN, 0, (size_t)n_xN * sizeof(double), h->stream));
    if ((rc = reserve(h, "pos0", dr.capt_full ? (size_t)h->n_sm * 4 * 1024 : 0, &m.pos0))) return rc;
    h->n_map = n_map; h->n_org = n_org;
    if (fresh && n_map) CK(cudaMemsetAsync(m.smap, 0, (size_t)n_map * sizeof(double), h->stream));
    if (fresh && n_org) CK(cudaMemsetAsync(m.star_origin, 0, (size_t)n_org * sizeof(double), h->stream));
  }
  if (n_local == 0 || (dr.count_sent && dr.n_packets_total == 0)) { CK(cudaEventRecord(h->ev0, h->stream)); CK(cudaEventRecord(h->ev1, h->stream)); h->launched = true; return MCB_OK; }
  if (h->kf_dark_stale) {
    std::vector<double> kd(h->host_kappa_factor);
    for (size_t i = 0; i < kd.size(); ++i) if (i < h->host_dark.size() && h->host_dark[i]) kd[i] = -fabs(kd[i]) - 0.0;   // (-0.0 keeps the flag when kappa_factor == 0)
    for (size_t i = 0; i < kd.size(); ++i) if (i < h->host_dark.size() && h->host_dark[i] && kd[i] == 0.0) kd[i] = -0.0;
    if ((rc = put(h, "kf_dark", kd.data(), kd.size(), &m.kf_dark))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    h->kf_dark_stale = false;
  }
  compute_smem_layout(h, r->p_lambda_in);
  rc = mcb_launch_mc(h, dr);
  if (rc) return rc;
  h->launched = true;
  return MCB_OK;
}

int mcfost_b200_set_overlap(mcb_handle* h, int n_sms_reserved, int n_sms_straggler) {
  if (!h) return MCB_ERR_BAD_ARG;
  if (n_sms_reserved == 0) { h->overlap_sms = 0; h->straggler_sms = 0; return MCB_OK; }
  if (n_sms_straggler <= 0) n_sms_straggler = n_sms_reserved;
  if (n_sms_reserved < 0 || n_sms_reserved > h->n_sm / 2 || n_sms_straggler > n_sms_reserved) return fail(h, MCB_ERR_BAD_ARG, "set_overlap: 0 <= straggler SMs <= reserved SMs <= half the SMs");
  h->overlap_sms = n_sms_reserved; h->straggler_sms = n_sms_straggler;
  return MCB_OK;
}

int mcfost_b200_sync(mcb_handle* h) {
  if (!h) return MCB_ERR_BAD_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_last_kernel_ms(mcb_handle* h, float* ms) {
  if (!h || !ms) return MCB_ERR_BAD_ARG;
  if (!h->launched) return fail(h, MCB_ERR_STATE, "no launch yet");
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return MCB_OK;
}

int mcfost_b200_tally_buffers(mcb_handle* h, void** d_f64, int64_t* n_f64, void** d_f32, int64_t* n_f32) {
  if (!h) return MCB_ERR_BAD_ARG;
  if (!h->n_tally) return fail(h, MCB_ERR_STATE, "no tallies allocated yet");
  if (d_f64) *d_f64 = h->m.tally;
  if (n_f64) *n_f64 = h->n_tally;
  if (d_f32) *d_f32 = h->n_xI ? h->m.xI : nullptr;
  if (n_f32) *n_f32 = h->n_xI;
  return MCB_OK;
}

int mcfost_b200_download(mcb_handle* h, const mcb_run_params* r, mcb_tallies* out) {
  if (!h || !out) return MCB_ERR_BAD_ARG;
  if (!h->n_tally) return fail(h, MCB_ERR_STATE, "download before launch");
  CK(cudaSetDevice(h->device));
  const DevModel& m = h->m;
  const TallyLayout& L = m.lay;
  auto get = [&](double* dst, int64_t off, int64_t n) -> cudaError_t {
    if (!dst || n <= 0) return cudaSuccess;
    return cudaMemcpyAsync(dst, m.tally + off, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  };
  CK(get(out->xKJ_abs, L.xKJ, m.n_cells));
  if (h->lay_xJ) CK(get(out->xJ_abs, L.xJ, (int64_t)m.n_cells * m.n_lambda));
  CK(get(out->n_phot_envoyes, L.n_env, m.n_lambda));
  double* sp[9] = {out->sed, out->sed_q, out->sed_u, out->sed_v, out->n_phot_sed, out->sed_star, out->sed_star_scat, out->sed_disk, out->sed_disk_scat};
  for (int a = 0; a < 9; ++a) CK(get(sp[a], L.sed + a * L.n_sed, L.n_sed));
  CK(get(out->stats, L.stats, 12));
  CK(get(out->E_abs_nRE, L.E_abs_nRE, 1));
  if (out->xN_abs && h->n_xN) CK(cudaMemcpyAsync(out->xN_abs, m.xN, (size_t)h->n_xN * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out->stokes_map && h->n_map) CK(cudaMemcpyAsync(out->stokes_map, m.smap, (size_t)h->n_map * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out->star_origin && h->n_org) CK(cudaMemcpyAsync(out->star_origin, m.star_origin, (size_t)m.n_lambda * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out->disk_origin && h->n_org) CK(cudaMemcpyAsync(out->disk_origin, m.disk_origin, (size_t)m.n_lambda * m.n_cells * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out->xT_ech_1grain && h->n_1g) CK(cudaMemcpyAsync(out->xT_ech_1grain, m.gr.xT_1g, (size_t)h->n_1g * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (out->xT_ech_1grain_nRE && h->n_1g_nRE) CK(cudaMemcpyAsync(out->xT_ech_1grain_nRE, m.gr.xT_1g_nRE, (size_t)h->n_1g_nRE * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (out->xT_ech) CK(cudaMemcpyAsync(out->xT_ech, m.xT_ech, (size_t)m.n_cells * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (out->xI_scatt && h->n_xI) CK(cudaMemcpyAsync(out->xI_scatt, m.xI, (size_t)h->n_xI * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (out->I_spec && h->n_Ispec) CK(cudaMemcpyAsync(out->I_spec, m.I_spec, (size_t)h->n_Ispec * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  if (out->I_spec_star && h->n_Ispec) CK(cudaMemcpyAsync(out->I_spec_star, m.I_spec_star, (size_t)m.n_cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  out->N_type_flux = h->n_type_flux;
  CK(cudaStreamSynchronize(h->stream));
  (void)r;
  return MCB_OK;
}

int mcfost_b200_run(mcb_handle* h, const mcb_run_params* r, mcb_tallies* out) {
  int rc = mcfost_b200_launch(h, r);
  if (rc) return rc;
  rc = mcfost_b200_sync(h);
  if (rc) return rc;
  if (out) rc = mcfost_b200_download(h, r, out);
  return rc;
}

}  // extern "C"

// ===========================================================================
// post-MC temperature solves (one cell / one (grain, cell) pair per thread), reference arithmetic order
// ===========================================================================
__global__ void temp_finale_kernel(const __grid_constant__ DevModel m, float* Tdust) {
  const int ic = blockIdx.x * blockDim.x + threadIdx.x;       // 0-based cell
  if (ic >= m.n_cells) return;
  const int pc = (m.p_n_cells != 1) ? ic : 0;
  const double* logQ = m.logQ + (size_t)m.n_T * pc;            // log_Qcool_minus_extra_heating(:, p_icell)
  const double Qheat = m.tally[m.lay.xKJ + ic] * m.L_packet_th / m.volume[ic];      // sum over id = the merged tally
  float Temp = m.T_min;
  if (!(Qheat < MCB_TINY_DP)) {
    const double log_Qheat = log(Qheat);
    if (!(log_Qheat < logQ[0])) {
      int Ti = m.xT_ech[ic];
      while ((logQ[Ti - 1] < log_Qheat) && (Ti < m.n_T)) ++Ti;
      // the photon loop's cached index can sit above the final one (it is an atomicMax of racing warps): step back
      while (Ti > 2 && !(logQ[Ti - 2] < log_Qheat)) --Ti;
      const double frac = (log_Qheat - logQ[Ti - 2]) / (logQ[Ti - 1] - logQ[Ti - 2]);
      Temp = (float)exp((double)logf(m.tab_Temp[Ti - 1]) * frac + (double)logf(m.tab_Temp[Ti - 2]) * (1.0 - frac));
    }
  }
  Tdust[ic] = Temp;
}

__global__ void temp_finale_nlte_kernel(const __grid_constant__ DevModel m, float* T1g) {
  const DevGrains& g = m.gr;
  const int nk = g.nLTE_e - g.nLTE_s + 1;
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)nk * m.n_cells) return;
  const int kk = (int)(t % nk), ic = (int)(t / nk);
  const int k = g.nLTE_s + kk;
  const int pk = (m.p_n_cells != 1) ? k : g.zone[k - 1];
  if (!(g.dd[(pk - 1) + (size_t)g.n_dens * ic] > MCB_TINY_DP)) { T1g[t] = 0.0f; return; }
  double J = 0.0;
  for (int l = 0; l < m.n_lambda; ++l) {
    const size_t cl = (size_t)ic + (size_t)m.n_cells * l;
    J = J + (double)g.C_abs_norm[(k - 1) + (size_t)g.n_grains_tot * l] * (m.tally[m.lay.xJ + cl] + g.J0[cl]);
  }
  J = J * m.L_packet_th / m.volume[ic];
  float Temp = m.T_min;
  if (!(J < MCB_TINY_DP)) {
    const double log_E = log(J);
    auto LE = [&](int Tt) { return g.logE[kk + (size_t)nk * (Tt - 1)]; };
    if (!(log_E < LE(1))) {
      int Ti = g.xT_1g[kk + (size_t)nk * ic];
      while ((LE(Ti) < log_E) && (Ti < m.n_T)) ++Ti;
      while (Ti > 2 && !(LE(Ti - 1) < log_E)) --Ti;
      const double T2 = m.tab_Temp[Ti - 1], T1 = m.tab_Temp[Ti - 2];
      const double frac = (log_E - LE(Ti - 1)) / (LE(Ti) - LE(Ti - 1));
      Temp = (float)exp(log(T2) * frac + log(T1) * (1.0 - frac));
    }
  }
  T1g[t] = Temp;
}

extern "C" {

int mcfost_b200_temp_finale(mcb_handle* h, float* Tdust) {
  if (!h || !Tdust) return MCB_ERR_BAD_ARG;
  if (!h->n_tally || !h->launched) return fail(h, MCB_ERR_STATE, "temp_finale before a photon-loop call");
  if (!h->m.logQ || !h->m.tab_Temp) return fail(h, MCB_ERR_BAD_ARG, "thermal tables missing");
  CK(cudaSetDevice(h->device));
  float* d = nullptr; int rc;
  if ((rc = reserve(h, "Tdust", (size_t)h->m.n_cells, &d))) return rc;
  temp_finale_kernel<<<(h->m.n_cells + 127) / 128, 128, 0, h->stream>>>(h->m, d);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(Tdust, d, (size_t)h->m.n_cells * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_temp_finale_nlte(mcb_handle* h, float* T1g) {
  if (!h || !T1g) return MCB_ERR_BAD_ARG;
  if (!h->n_tally || !h->launched) return fail(h, MCB_ERR_STATE, "temp_finale_nlte before a photon-loop call");
  if (!h->has_gr || !h->n_1g || !h->lay_xJ) return fail(h, MCB_ERR_STATE, "temp_finale_nlte needs a call with lRE_nLTE and xJ_abs");
  const mcb_grains& gh = h->gr_host;
  if (!gh.C_abs_norm || !gh.J0 || !gh.log_E_em_1grain || (h->m.p_n_cells == 1 && !gh.grain_zone)) return fail(h, MCB_ERR_BAD_ARG, "nLTE tables missing");
  CK(cudaSetDevice(h->device));
  float* d = nullptr; int rc;
  if ((rc = reserve(h, "T1g", (size_t)h->n_1g, &d))) return rc;
  temp_finale_nlte_kernel<<<(unsigned)((h->n_1g + 127) / 128), 128, 0, h->stream>>>(h->m, d);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(T1g, d, (size_t)h->n_1g * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

}  // extern "C"

// ===========================================================================
// deterministic sub-kernels: one ray per thread
// ===========================================================================
template <class G>
__global__ void cross_cell_kernel(const __grid_constant__ DevModel m, int64_t n, const double* x0, const double* y0, const double* z0,
                                  const double* u, const double* v, const double* w, const int* icell, const int* prev,
                                  double* x1, double* y1, double* z1, int* next_cell, double* l, double* lc, double* lv) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c, p, nx;
  cell_of_id(m, icell[i], c);
  if (prev[i] != 0) cell_of_id(m, prev[i], p); else null_cell(p);
  DirInv d = dir_invariants(u[i], v[i], w[i]);
  double a, b, cc, lcon, lvoid;
  double ll = G::cross(m, d, x0[i], y0[i], z0[i], u[i], v[i], w[i], c, p, a, b, cc, nx, lcon, lvoid);
  x1[i] = a; y1[i] = b; z1[i] = cc; next_cell[i] = id_of_cell(m, nx); l[i] = ll; lc[i] = lcon; lv[i] = lvoid;
}

template <class G>
__global__ void index_cell_kernel(const __grid_constant__ DevModel m, int64_t n, const double* x, const double* y, const double* z, int* icell) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  icell[i] = id_of_cell(m, G::index(m, x[i], y[i], z[i]));
}

template <class G>
__global__ void move_to_grid_kernel(const __grid_constant__ DevModel m, int64_t n, double* x, double* y, double* z,
                                    const double* u, const double* v, const double* w, int* icell, int* lintersect) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c; null_cell(c);
  double a = x[i], b = y[i], cc = z[i];
  bool ok = G::move_to_grid(m, a, b, cc, u[i], v[i], w[i], c);
  lintersect[i] = ok;
  icell[i] = ok ? id_of_cell(m, c) : 0;
  if (ok) { x[i] = a; y[i] = b; z[i] = cc; }
}

// distance_to_closest_wall (grid.f90 procedure pointer)
template <class G>
__global__ void closest_wall_kernel(const __grid_constant__ DevModel m, int64_t n, const int* icell, const double* x, const double* y, const double* z, double* s) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c; cell_of_id(m, icell[i], c);
  s[i] = G::closest_wall(m, c, x[i], y[i], z[i]);
}

// optical_depth.f90:248-324
template <class G>
__global__ void optical_length_tot_kernel(const __grid_constant__ DevModel m, int64_t n, int lambda, const double* x, const double* y, const double* z,
                                          const double* u, const double* v, const double* w, const int* icell,
                                          double* tau_out, double* lmin_out, double* lmax_out, int* nsteps) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c0, c_prev, c1;
  cell_of_id(m, icell[i], c0); null_cell(c_prev);
  const double uu = u[i], vv = v[i], ww = w[i];
  DirInv d = dir_invariants(uu, vv, ww);
  double x0 = x[i], y0 = y[i], z0 = z[i];
  double tau_tot = 0.0, lmin = 0.0, ltot = 0.0;
  int ns = 0;
  const bool variable_dust = m.p_n_cells != 1;
  for (;;) {
    if (G::test_exit(m, c0, x0, y0, z0)) break;
    const int idx = tally_index(m, c0);
    double opacity = 0.0;
    if (idx >= 0) {
      const int p_icell = variable_dust ? idx + 1 : 1;
      opacity = __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
    }
    double x1, y1, z1, lcon, lvoid;
    double l = G::cross(m, d, x0, y0, z0, uu, vv, ww, c0, c_prev, x1, y1, z1, c1, lcon, lvoid);
    ++ns;
    tau_tot = tau_tot + lcon * opacity;
    ltot = ltot + l;
    if (tau_tot < MCB_TINY_REAL) lmin = ltot;
    c_prev = c0; c0 = c1; x0 = x1; y0 = y1; z0 = z1;
    if (ns > 100000000) break;
  }
  tau_out[i] = (double)(float)tau_tot;      // tau_tot_out is `real`
  lmin_out[i] = lmin; lmax_out[i] = ltot;
  if (nsteps) nsteps[i] = ns;
}

// optical_depth.f90:328-415 compute_column: one thread per (cell, direction)
template <class G>
__global__ void compute_column_kernel(const __grid_constant__ DevModel m, int lambda, const double* factor, const double* cx, const double* cy,
                                      const double* cz, float* column) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= 4 * (int64_t)m.n_cells) return;
  const int icell = (int)(i % m.n_cells) + 1, direction = (int)(i / m.n_cells) + 1;
  double x0 = cx[icell - 1], y0 = cy[icell - 1], z0 = cz[icell - 1], uu, vv, ww;
  if (direction == 1) { const double norm = 1.0 / sqrt(x0 * x0 + y0 * y0 + z0 * z0); uu = -x0 * norm; vv = -y0 * norm; ww = -z0 * norm; }
  else if (direction == 2) { uu = 0.0; vv = 0.0; ww = 1.0; }
  else if (direction == 3) { uu = 0.0; vv = 0.0; ww = -1.0; }
  else { uu = x0; vv = y0; ww = 0.0; const double norm = 1.0 / sqrt(uu * uu + vv * vv); uu = uu * norm; vv = vv * norm; }
  typename G::CellT c0, c_prev, c1;
  cell_of_id(m, icell, c0); null_cell(c_prev);
  const DirInv d = dir_invariants(uu, vv, ww);
  const bool variable_dust = m.p_n_cells != 1;
  double sum = 0.0;
  for (int ns = 0; ns < 100000000; ++ns) {
    if (G::test_exit(m, c0, x0, y0, z0)) break;
    double x1, y1, z1, lcon, lvoid;
    G::cross(m, d, x0, y0, z0, uu, vv, ww, c0, c_prev, x1, y1, z1, c1, lcon, lvoid);
    const int idx = tally_index(m, c0);
    if (idx >= 0) {
      double f;
      if (factor) f = __ldg(factor + idx);
      else {
        const int p_icell = variable_dust ? idx + 1 : 1;
        f = __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
      }
      sum = sum + lcon * f;
    }
    c_prev = c0; c0 = c1; x0 = x1; y0 = y1; z0 = z1;
  }
  column[i] = (float)sum;
}

// ---- define_dark_zone (optical_depth.f90:1425-1651) in ONE block -----------------------------------------------
// The columns are a sequential chain (the rays of column i bounce off the cells that columns < i made dark), the
// <= nz x 11 rays of a column are independent: one thread per ray, a block-wide maximum picks the row the reference's
// nested loops would stop at, then the column is marked and the block moves on.  `real` sums as in the Fortran;
// cos / sin of the `real` angles come from the host's libm (the compiler's own single-precision cos / sin).
template <class G>
__device__ bool dark_walk_exits(const DevModel& m, const int* dark, int lambda, int icell, double x0, double y0, double z0,
                                double uu, double vv, double ww, float tau) {
  typename G::CellT c0, c_old, c1;
  cell_of_id(m, icell, c0); null_cell(c_old);
  const DirInv d = dir_invariants(uu, vv, ww);
  double extr = (double)tau;
  const int i_star_hit = intersect_stars(m, x0, y0, z0, uu, vv, ww);
  const bool variable_dust = m.p_n_cells != 1;
  for (int ns = 0; ns < 100000000; ++ns) {
    if (G::test_exit(m, c0, x0, y0, z0)) return true;
    if (i_star_hit > 0) {
      typename G::CellT cs; cell_of_id(m, m.star_icell[i_star_hit - 1], cs);
      if (same_cell(c0, cs)) return true;
    }
    const int idx = tally_index(m, c0);
    double opacity = 0.0;
    if (idx >= 0) {
      const int p_icell = variable_dust ? idx + 1 : 1;
      opacity = __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
      if (__ldcg(dark + idx)) return false;
    }
    double x1, y1, z1, lcon, lvoid;
    G::cross(m, d, x0, y0, z0, uu, vv, ww, c0, c_old, x1, y1, z1, c1, lcon, lvoid);
    const double tau_c = lcon * opacity;
    if (tau_c > extr) return false;
    extr = extr - tau_c;
    c_old = c0; x0 = x1; y0 = y1; z0 = z1; c0 = c1;
  }
  return false;
}

template <class G>
__global__ void __launch_bounds__(1024, 1)
define_dark_zone_kernel(const __grid_constant__ DevModel m, int lambda, float tau_max, const double* r_grid, const double* z_grid,
                        const float* cs_ang, const float* cs_phi, const double* dust_sum, int n_regions, const int* iRmin,
                        const int* iRmax, int* dark, int* ri_in, int* ri_out, int* zj_sup, int* zj_inf, int* l_is_dark) {
  constexpr int NA = 11;
  __shared__ int s_top, s_flag;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int n_rad = m.n_rad, nz = m.nz, n_az = m.n_az;
  const bool l3D = m.l3D != 0, lcyl = m.kind == MCB_GRID_CYL;
  const bool variable_dust = m.p_n_cells != 1;
  auto idx_of = [&](int i, int j, int pk) { Cell c; c.ri = i; c.zj = j; c.k = pk; return real_index(m, c); };
  auto kap = [&](int idx) {
    const int p_icell = variable_dust ? idx + 1 : 1;
    return __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
  };
  auto ZS = [&](int i, int pk) -> int& { return zj_sup[(size_t)(i - 1) + (size_t)n_rad * (pk - 1)]; };
  auto ZI = [&](int i, int pk) -> int& { return zj_inf[(size_t)(i - 1) + (size_t)n_rad * (pk - 1)]; };
  if (tid == 0) { s_top = 0; s_flag = 0; }
  for (int i = tid; i < m.n_cells; i += nt) dark[i] = 0;
  // ---- steps 1 - 3.5: where the optical depth exceeds tau_max radially (both ways) and vertically
  for (int pk = 1; pk <= n_az; ++pk) {
    if (tid == 0) {
      int rin = n_rad, rout = 1;
      float total_sum = 0.0f;
      for (int i = 1; i <= n_rad; ++i) {
        total_sum = (float)((double)total_sum + kap(idx_of(i, 1, pk)) * (m.r_lim[i] - m.r_lim[i - 1]));
        if (total_sum > tau_max) { rin = i; break; }
      }
      total_sum = 0.0f;
      for (int i = n_rad; i >= 1; --i) {
        total_sum = (float)((double)total_sum + kap(idx_of(i, 1, pk)) * (m.r_lim[i] - m.r_lim[i - 1]));
        if (total_sum > tau_max) { rout = i; break; }
      }
      if (rout == n_rad) rout = n_rad - 1;
      ri_in[pk - 1] = rin; ri_out[pk - 1] = rout;
    }
    __syncthreads();
    const int rin = ri_in[pk - 1], rout = ri_out[pk - 1];
    if (lcyl) {
      for (int i = rin + tid; i <= rout; i += nt) {
        float total_sum = 0.0f;
        for (int j = nz; j >= 1; --j) {
          total_sum = (float)((double)total_sum + kap(idx_of(i, j, pk)) * (z_lim<false>(m, i, j + 1) - z_lim<false>(m, i, j)));
          if (total_sum > tau_max) { ZS(i, pk) = j; break; }
        }
        if (l3D) {
          total_sum = 0.0f;
          for (int j = -nz; j <= -1; ++j) {
            total_sum = (float)((double)total_sum + kap(idx_of(i, j, pk)) * (z_lim<false>(m, i, -j + 1) - z_lim<false>(m, i, -j)));
            if (total_sum > tau_max) { ZI(i, pk) = j; break; }
          }
        }
      }
    } else {
      for (int i = 1 + tid; i <= n_rad; i += nt) ZS(i, pk) = nz;
    }
    __syncthreads();
  }
  __threadfence_block();
  __syncthreads();
  // ---- step 4: 11 rays from the centre of every candidate cell, column after column
  for (int pk = 1; pk <= n_az; ++pk) {
    const float cphi = l3D ? cs_phi[2 * (pk - 1)] : 1.0f, sphi = l3D ? cs_phi[2 * (pk - 1) + 1] : 0.0f;
    const int rin = max(ri_in[pk - 1], 2), rout = ri_out[pk - 1];
    for (int half = 0; half < (l3D ? 2 : 1); ++half) {
      for (int i = rin; i <= rout; ++i) {
        // upper half: rows j = zj_sup .. 1 (descending); lower half (3D): rows j = zj_inf .. -1 (ascending)
        const int jfirst = half == 0 ? ZS(i, pk) : max(ZI(i, pk), -nz);
        const int nrows = half == 0 ? jfirst : (ZI(i, pk) == 0 ? 0 : -jfirst);
        for (int t = tid; t < nrows * NA; t += nt) {
          const int row = t / NA, n = t % NA + 1;
          const int j = half == 0 ? jfirst - row : jfirst + row;
          const int idx = idx_of(i, j, pk);
          double x0, y0, z0;
          if (l3D) {
            const float r0 = (float)r_grid[idx];
            x0 = (double)(r0 * cphi); y0 = (double)(r0 * sphi); z0 = half == 0 ? z_grid[idx] : -z_grid[idx];
          } else { x0 = r_grid[idx]; y0 = 0.0; z0 = z_grid[idx]; }
          const double u0 = (double)cs_ang[2 * (n - 1)], w0 = (double)cs_ang[2 * (n - 1) + 1];
          if (!dark_walk_exits<G>(m, dark, lambda, idx + 1, x0, y0, z0, u0, 0.0, w0, tau_max)) {
            if (half == 0) atomicMax(&s_top, j); else s_flag = 1;
          }
        }
        __syncthreads();
        const int top = s_top;
        if (half == 0 && top > 0) {
          for (int jj = 1 + tid; jj <= top; jj += nt) __stcg(dark + idx_of(i, jj, pk), 1);
          if (!l3D && tid == 0) s_flag = 1;        // (the 3D upper-half loop does not set l_is_dark_zone, :1575-1582)
        }
        __threadfence_block();
        __syncthreads();
        if (tid == 0) s_top = 0;
        __syncthreads();
      }
    }
  }
  // ---- tidy up (:1620-1648)
  for (int pk = 1 + tid; pk <= n_az; pk += nt) {
    for (int i = 1; i <= ri_in[pk - 1] - 1; ++i) ZS(i, pk) = ZS(ri_in[pk - 1], pk);
    for (int i = ri_out[pk - 1] + 1; i <= n_rad; ++i) ZS(i, pk) = ZS(ri_out[pk - 1], pk);
    if (l3D) {
      for (int i = 1; i <= ri_in[pk - 1] - 1; ++i) ZI(i, pk) = ZI(ri_in[pk - 1], pk);
      for (int i = ri_out[pk - 1] + 1; i <= n_rad; ++i) ZI(i, pk) = ZI(ri_out[pk - 1], pk);
    }
  }
  if (dust_sum) for (int i = tid; i < m.n_cells; i += nt) if (dust_sum[i] < MCB_TINY_REAL) dark[i] = 0;
  __syncthreads();
  for (int q = 0; q < n_regions; ++q)
    for (int j = 1 + tid; j <= nz; j += nt) { dark[idx_of(iRmin[q], j, 1)] = 0; dark[idx_of(iRmax[q], j, 1)] = 0; }
  if (tid == 0) *l_is_dark = s_flag;
}

// ---- init_reemission (thermal_emission.f90:404-644) on the device ---------------------------------------------
// B(lambda, T) and dB/dT(lambda, T) with the reference's constants (`thermal_const` is a `real` parameter, 1.e-6 and 500.0
// are `real` literals), then one thread per (T, row): the cooling sum and the emission CDF of a row of absorption
// coefficients -- the LTE cells (kappa_abs_LTE(p_icell, :), CDF from lambda = 1) or single grains (C_abs_norm(k, :), `real`,
// CDF starting at 0 for lambda = 1).
__global__ void planck_tables_kernel(int n_lambda, int n_T, const double* tab_lambda, const double* tab_delta_lambda, const float* tab_Temp,
                                     double* B, double* dB) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_lambda * n_T) return;
  const int lambda = q % n_lambda, t = q / n_lambda;
  const float thermal_const = (float)(299792458.0 * 6.626070040e-34 / 1.38064852e-23);
  const double Temp = (double)tab_Temp[t];
  const double cst = (double)thermal_const / Temp;
  const double wl = tab_lambda[lambda] * (double)1.e-6f, delta_wl = tab_delta_lambda[lambda] * (double)1.e-6f;
  const double cst_wl = cst / wl;
  double b = 0.0, db = 0.0;
  if (cst_wl < 500.0) {
    const double coeff_exp = exp(cst_wl);
    const double wl2 = wl * wl, wl5 = (wl2 * wl2) * wl;
    b = 1.0 / (wl5 * (coeff_exp - 1.0)) * delta_wl;
    db = b * cst_wl * coeff_exp / (coeff_exp - 1.0);
  }
  B[q] = b; dB[q] = db;
}
template <bool GRAINS>
__global__ void init_reemission_rows_kernel(int n_lambda, int n_T, int n_rows, const double* a_cells, int stride_cells, const float* a_grains,
                                            int stride_grains, int k0, const double* B, const double* dB, double* logQ, double* E_em, double* cdf) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= (int64_t)n_T * n_rows) return;
  // cells: logQ(T, icell), cdf(lambda, T, icell);  grains: logE(k, T), cdf(lambda, k, T)
  const int t = GRAINS ? (int)(q / n_rows) : (int)(q % n_T), row = GRAINS ? (int)(q % n_rows) : (int)(q / n_T);
  auto a = [&](int l) { return GRAINS ? (double)a_grains[(size_t)(k0 + row) + (size_t)stride_grains * l] : a_cells[(size_t)row + (size_t)stride_cells * l]; };
  const double cst_E = 2.0 * 6.626070040e-34 * (299792458.0 * 299792458.0) * (4.0 * MCB_PI);
  double integ = 0.0, integ0 = 0.0;
  for (int l = 0; l < n_lambda; ++l) { integ = integ + a(l) * B[l + (size_t)n_lambda * t]; if (!GRAINS) integ0 = integ0 + a(l) * B[l]; }
  const size_t o = GRAINS ? (size_t)row + (size_t)n_rows * t : (size_t)t + (size_t)n_T * row;
  if (GRAINS) {
    logQ[o] = (integ > MCB_TINY_DP) ? log(integ * cst_E) : -1000.0;
    if (E_em) E_em[o] = integ * cst_E;
  } else {
    const double qc = integ * cst_E - integ0 * cst_E;      // Qcool - Qcool0 (no extra heating: the cloud at T_min, :467-470)
    logQ[o] = (qc > MCB_TINY_DP) ? log(qc) : -1000.0;
  }
  double* c = cdf + (size_t)n_lambda * o;
  double run = 0.0;
  for (int l = GRAINS ? 1 : 0; l < n_lambda; ++l) run = run + a(l) * dB[l + (size_t)n_lambda * t];
  const double tot = run;
  run = 0.0;
  if (tot > MCB_TINY_DP) {
    if (GRAINS) c[0] = 0.0 / tot;
    for (int l = GRAINS ? 1 : 0; l < n_lambda; ++l) { run = run + a(l) * dB[l + (size_t)n_lambda * t]; c[l] = run / tot; }
  } else for (int l = 0; l < n_lambda; ++l) c[l] = 0.0;
}

// ---- ray-tracing method 1: source function and formal solution (dust_ray_tracing.f90:636-708,1458-1485; optical_depth.f90:1327-1421)
// eps_dust1(k, psup, itype, icell) with the storage extents of the xI_scatt tally (N_AZ_RT x 2); on a 3D grid only (1, 1) is used.
__global__ void init_dust_source_fct1_kernel(const __grid_constant__ DevModel m, int lambda, int iRT, int n_RT, double photon_energy,
                                             const double* J_th, const float* xI, int n_az_rt, int n_theta_rt, int ntf, int n_stokes, int pola,
                                             int contrib, double* eps) {
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= (int64_t)N_AZ_RT * 2 * m.n_cells) return;
  const int k = (int)(q % N_AZ_RT) + 1, psup = (int)((q / N_AZ_RT) % 2) + 1, idx = (int)(q / (2 * N_AZ_RT));
  double* e = eps + (size_t)(k - 1) + (size_t)N_AZ_RT * ((size_t)(psup - 1) + 2 * ((size_t)ntf * (size_t)idx));
  const size_t stride = (size_t)N_AZ_RT * 2;
  for (int it = 0; it < ntf; ++it) e[stride * it] = 0.0;
  if (k > n_az_rt || psup > n_theta_rt) return;
  const int p_icell = (m.p_n_cells != 1) ? idx + 1 : 1;
  const double factor = photon_energy / m.volume[idx] * n_az_rt * n_theta_rt;
  const double kappa_ext = m.kappa[(p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)] * m.kappa_factor[idx];
  const double kappa_sca = kappa_ext * m.albedo[(p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)];
  if (!(kappa_ext > MCB_TINY_DP)) return;
  const float* xi = xI + (size_t)(k - 1) + (size_t)N_AZ_RT * ((size_t)(psup - 1) + 2 * ((size_t)ntf * ((size_t)(iRT - 1) + (size_t)n_RT * (size_t)idx)));
  auto I_scatt = [&](int itype) { return (double)xi[stride * (itype - 1)] * factor * kappa_sca; };
  e[0] = (I_scatt(1) + J_th[idx]) / kappa_ext;
  if (pola) for (int it = 2; it <= 4; ++it) e[stride * (it - 1)] = I_scatt(it) / kappa_ext;
  if (contrib) {
    e[stride * (n_stokes + 1)] = I_scatt(n_stokes + 2) / kappa_ext;
    e[stride * (n_stokes + 2)] = J_th[idx] / kappa_ext;
    e[stride * (n_stokes + 3)] = I_scatt(n_stokes + 4) / kappa_ext;
  }
}

template <class G>
__global__ void integ_ray_dust_kernel(const __grid_constant__ DevModel m, int64_t n, int lambda, const double* x, const double* y, const double* z,
                                      const double* u, const double* v, const double* w, const int* icell, float tau_dark_zone_obs,
                                      const double* eps, int n_az_rt, int ntf, double* out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c0, c_prev, c1;
  cell_of_id(m, icell[i], c0); null_cell(c_prev);
  const double uu = u[i], vv = v[i], ww = w[i];
  const DirInv d = dir_invariants(uu, vv, ww);
  double x0 = x[i], y0 = y[i], z0 = z[i];
  const int i_star_hit = intersect_stars(m, x0, y0, z0, uu, vv, ww);
  const bool variable_dust = m.p_n_cells != 1;
  double acc[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) acc[it] = 0.0;
  double tau = 0.0;
  const size_t stride = (size_t)N_AZ_RT * 2;
  for (int ns = 0; ns < 100000000; ++ns) {
    if (G::test_exit(m, c0, x0, y0, z0)) break;
    if (i_star_hit > 0) {
      typename G::CellT cs; cell_of_id(m, m.star_icell[i_star_hit - 1], cs);
      if (same_cell(c0, cs)) break;
    }
    double x1, y1, z1, lcon, lvoid;
    G::cross(m, d, x0, y0, z0, uu, vv, ww, c0, c_prev, x1, y1, z1, c1, lcon, lvoid);      // (previous_cell = 0 at every step, :1388)
    const int idx = tally_index(m, c0);
    if (idx >= 0) {
      const int p_icell = variable_dust ? idx + 1 : 1;
      const double dtau = lcon * __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
      const double xm = 0.5 * (x0 + x1), ym = 0.5 * (y0 + y1), zm = 0.5 * (z0 + z1);
      int k = 1, psup = 1;
      if (!m.l3D) {
        psup = (zm > 0.0) ? 1 : 2;
        const double phi_pos = atan2(xm, ym);
        k = (int)floor(fmodulo(phi_pos, MCB_TWO_PI) / MCB_TWO_PI * n_az_rt) + 1;
        if (k > n_az_rt) k = n_az_rt;
      }
      const double wgt = exp(-tau) * (1.0 - exp(-dtau));
      const double* e = eps + (size_t)(k - 1) + (size_t)N_AZ_RT * ((size_t)(psup - 1) + 2 * ((size_t)ntf * (size_t)idx));
#pragma unroll
      for (int it = 0; it < 8; ++it) if (it < ntf) acc[it] = acc[it] + wgt * __ldg(e + stride * it);
      tau = tau + dtau;
      if (tau > (double)tau_dark_zone_obs) break;
    }
    x0 = x1; y0 = y1; z0 = z1; c0 = c1;
  }
  for (int it = 0; it < ntf; ++it) out[(size_t)ntf * i + it] = acc[it];
}

// ---- repartition_energie (thermal_emission.f90:1771-1949), LTE case, on the device ------------------------------
// E_cell(icell) = 4 kappa_abs_LTE kappa_factor volume / (wl^5 (exp(hc / (k T wl)) - 1)) from Tdust (`real`); its sum is
// E_disk(lambda), its cumulative sum (weighted by weight_proba_emission when given) prob_E_cell(0:n_cells, lambda).
__global__ void emission_cells_kernel(const __grid_constant__ DevModel m, int lambda, double wl, const float* Tdust, const double* weight,
                                      double* E_cell, double* E_corr) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m.n_cells) return;
  const float thermal_const = (float)(299792458.0 * 6.626070040e-34 / 1.38064852e-23);
  const double cst_wl_max = (double)(logf(FLT_MAX) - 1.0e-4f);
  const double wl2 = wl * wl, wl5 = (wl2 * wl2) * wl;
  double E = 0.0;
  if (!(m.dark && m.dark[idx])) {
    const double Temp = (double)Tdust[idx];
    if (!(Temp < MCB_TINY_REAL)) {
      const double cst_wl = (double)thermal_const / (Temp * wl);
      if (cst_wl < cst_wl_max) {
        const int p_icell = (m.p_n_cells != 1) ? idx + 1 : 1;
        E = 4.0 * m.kappa_abs[(p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)] * m.kappa_factor[idx] * m.volume[idx] / (wl5 * (exp(cst_wl) - 1.0));
      }
    }
  }
  E_cell[idx] = E;
  E_corr[idx] = weight ? E * weight[idx] : E;
}
// prob(0:n_cells) of one wavelength from the inclusive sums cum(1:n_cells) (:1923-1941)
__global__ void normalise_cdf_kernel(const double* cum, int n_cells, double* prob) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_cells) return;
  const double tot = cum[n_cells - 1];
  prob[i] = (tot > MCB_TINY_DP && i > 0) ? cum[i - 1] / tot : 0.0;
}

// optical_depth.f90:21-182 with Stokes = 0 (no tallies)
template <class G>
__global__ void physical_length_kernel(const __grid_constant__ DevModel m, int64_t n, int lambda, double* x, double* y, double* z,
                                       double* u, double* v, double* w, int* icell, const float* tau, float* ltot_out,
                                       int* flag_sortie, int* alive_out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename G::CellT c0, c_old, c1;
  cell_of_id(m, icell[i], c0); null_cell(c_old);
  int id_old = 0;
  double uu = u[i], vv = v[i], ww = w[i];
  DirInv d = dir_invariants(uu, vv, ww);
  double x0 = x[i], y0 = y[i], z0 = z[i], xo = x0, yo = y0, zo = z0;
  double extr = (double)tau[i];
  float ltot = 0.0f;
  const int i_star_hit = intersect_stars(m, x0, y0, z0, uu, vv, ww);
  const bool variable_dust = m.p_n_cells != 1;
  int sortie = 0, alive = 1;
  for (;;) {
    if (G::test_exit(m, c0, x0, y0, z0)) { sortie = 1; break; }
    if (i_star_hit > 0) {
      typename G::CellT cs; cell_of_id(m, m.star_icell[i_star_hit - 1], cs);
      if (same_cell(c0, cs)) { alive = 0; sortie = 1; break; }
    }
    const int idx = tally_index(m, c0);
    double opacity = 0.0;
    if (idx >= 0) {
      const int p_icell = variable_dust ? idx + 1 : 1;
      opacity = __ldg(m.kappa + (p_icell - 1) + (size_t)m.p_n_cells * (lambda - 1)) * __ldg(m.kappa_factor + idx);
      if (__ldg(m.dark + idx)) {
        u[i] = -uu; v[i] = -vv; w[i] = -ww;
        icell[i] = id_old; x[i] = xo; y[i] = yo; z[i] = zo;
        ltot_out[i] = ltot; flag_sortie[i] = 0; alive_out[i] = alive;
        return;
      }
    }
    double x1, y1, z1, lcon, lvoid;
    double l = G::cross(m, d, x0, y0, z0, uu, vv, ww, c0, c_old, x1, y1, z1, c1, lcon, lvoid);
    const double tau_c = lcon * opacity;
    if (tau_c > extr) {
      lcon = lcon * (extr / tau_c);
      l = lvoid + lcon;
      ltot = (float)((double)ltot + l);
      double xf = x0 + l * uu, yf = y0 + l * vv, zf = z0 + l * ww;
      typename G::CellT cf = c0;
      if (!G::is_vor && m.l3D && m.kind == 1) cf = G::index(m, xf, yf, zf);
      x[i] = xf; y[i] = yf; z[i] = zf; icell[i] = id_of_cell(m, cf);
      ltot_out[i] = ltot; flag_sortie[i] = 0; alive_out[i] = alive;
      return;
    }
    extr = extr - tau_c;
    ltot = (float)((double)ltot + l);
    id_old = id_of_cell(m, c0);
    xo = x0; yo = y0; zo = z0; c_old = c0;
    x0 = x1; y0 = y1; z0 = z1; c0 = c1;
  }
  // exit: the reference leaves xio,yio,zio,icell untouched
  ltot_out[i] = ltot; flag_sortie[i] = sortie; alive_out[i] = alive;
}

// ---- host wrappers (host pointers in, host pointers out) --------------------
namespace {
struct Scratch {
  mcb_handle* h;
  std::vector<void*> d;
  ~Scratch() { for (void* p : d) cudaFree(p); }
  template <class T> T* in(const T* src, int64_t n) {
    T* p = nullptr;
    if (cudaMalloc(&p, (size_t)(n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    d.push_back(p);
    if (src) cudaMemcpyAsync(p, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice, h->stream);
    return p;
  }
  template <class T> void out(T* dst, const T* src, int64_t n) {
    if (dst) cudaMemcpyAsync(dst, src, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, h->stream);
  }
};
}  // namespace

#define DISPATCH(KERNEL, ...)                                                              \
  switch (h->gk) {                                                                         \
    case GK_CYL2D: KERNEL<GeomCyl<false>><<<nb, 128, 0, h->stream>>>(__VA_ARGS__); break;  \
    case GK_CYL3D: KERNEL<GeomCyl<true>><<<nb, 128, 0, h->stream>>>(__VA_ARGS__); break;   \
    case GK_SPH2D: KERNEL<GeomSph<false>><<<nb, 128, 0, h->stream>>>(__VA_ARGS__); break;  \
    case GK_SPH3D: KERNEL<GeomSph<true>><<<nb, 128, 0, h->stream>>>(__VA_ARGS__); break;   \
    case GK_VOR:   KERNEL<GeomVor><<<nb, 128, 0, h->stream>>>(__VA_ARGS__); break;         \
  }

extern "C" {

int mcfost_b200_cross_cell(mcb_handle* h, int64_t n, const double* x0, const double* y0, const double* z0,
                           const double* u, const double* v, const double* w, const int32_t* icell, const int32_t* previous_cell,
                           double* x1, double* y1, double* z1, int32_t* next_cell, double* l, double* l_contrib, double* l_void_before) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "cross_cell before upload_grid");
  if (n == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  std::vector<int32_t> zero;
  if (!previous_cell) { zero.assign((size_t)n, 0); previous_cell = zero.data(); }
  const double *dx0 = s.in(x0, n), *dy0 = s.in(y0, n), *dz0 = s.in(z0, n), *du = s.in(u, n), *dv = s.in(v, n), *dw = s.in(w, n);
  const int *dic = s.in(icell, n), *dpr = s.in(previous_cell, n);
  double *dx1 = s.in<double>(nullptr, n), *dy1 = s.in<double>(nullptr, n), *dz1 = s.in<double>(nullptr, n);
  double *dl = s.in<double>(nullptr, n), *dlc = s.in<double>(nullptr, n), *dlv = s.in<double>(nullptr, n);
  int* dnx = s.in<int>(nullptr, n);
  if (!dnx) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(cross_cell_kernel, h->m, n, dx0, dy0, dz0, du, dv, dw, dic, dpr, dx1, dy1, dz1, dnx, dl, dlc, dlv);
  CK(cudaGetLastError());
  s.out(x1, dx1, n); s.out(y1, dy1, n); s.out(z1, dz1, n); s.out(next_cell, dnx, n); s.out(l, dl, n); s.out(l_contrib, dlc, n); s.out(l_void_before, dlv, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_index_cell(mcb_handle* h, int64_t n, const double* x, const double* y, const double* z, int32_t* icell) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "index_cell before upload_grid");
  if (n == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  const double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n);
  int* dic = s.in<int>(nullptr, n);
  if (!dic) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(index_cell_kernel, h->m, n, dx, dy, dz, dic);
  CK(cudaGetLastError());
  s.out(icell, dic, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_distance_to_closest_wall(mcb_handle* h, int64_t n, const int32_t* icell, const double* x, const double* y, const double* z, double* s_out) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "distance_to_closest_wall before upload_grid");
  if (n == 0) return MCB_OK;
  const DevModel& m = h->m;
  if (h->gk != GK_VOR && !m.r_lim) return fail(h, MCB_ERR_BAD_ARG, "r_lim missing");
  if ((h->gk == GK_SPH2D || h->gk == GK_SPH3D) && !m.w_lim) return fail(h, MCB_ERR_BAD_ARG, "w_lim missing");
  if ((h->gk == GK_CYL3D || h->gk == GK_SPH3D) && (!m.sin_phi_lim || !m.cos_phi_lim)) return fail(h, MCB_ERR_BAD_ARG, "sin_phi_lim / cos_phi_lim missing");
  for (int64_t i = 0; i < n; ++i) if (icell[i] < 1 || icell[i] > m.n_cells) return fail(h, MCB_ERR_BAD_ARG, "distance_to_closest_wall: icell is not a real cell");
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  const int* dic = s.in(icell, n);
  const double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n);
  double* ds = s.in<double>(nullptr, n);
  if (!ds) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(closest_wall_kernel, h->m, n, dic, dx, dy, dz, ds);
  CK(cudaGetLastError());
  s.out(s_out, ds, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_move_to_grid(mcb_handle* h, int64_t n, double* x, double* y, double* z, const double* u, const double* v, const double* w,
                             int32_t* icell, int32_t* lintersect) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "move_to_grid before upload_grid");
  if (n == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n);
  const double *du = s.in(u, n), *dv = s.in(v, n), *dw = s.in(w, n);
  int *dic = s.in<int>(nullptr, n), *dli = s.in<int>(nullptr, n);
  if (!dli) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(move_to_grid_kernel, h->m, n, dx, dy, dz, du, dv, dw, dic, dli);
  CK(cudaGetLastError());
  s.out(x, dx, n); s.out(y, dy, n); s.out(z, dz, n); s.out(icell, dic, n); s.out(lintersect, dli, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_optical_length_tot(mcb_handle* h, int64_t n, int32_t lambda, const double* x, const double* y, const double* z,
                                   const double* u, const double* v, const double* w, const int32_t* icell,
                                   double* tau_tot, double* lmin, double* lmax, int32_t* n_steps) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "optical_length_tot before upload_grid/opacity");
  if (lambda < 1 || lambda > h->m.n_lambda) return fail(h, MCB_ERR_BAD_ARG, "lambda out of range");
  if (n == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  const double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n), *du = s.in(u, n), *dv = s.in(v, n), *dw = s.in(w, n);
  const int* dic = s.in(icell, n);
  double *dt = s.in<double>(nullptr, n), *dmin = s.in<double>(nullptr, n), *dmax = s.in<double>(nullptr, n);
  int* dns = s.in<int>(nullptr, n);
  if (!dns) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(optical_length_tot_kernel, h->m, n, lambda, dx, dy, dz, du, dv, dw, dic, dt, dmin, dmax, dns);
  CK(cudaGetLastError());
  s.out(tau_tot, dt, n); s.out(lmin, dmin, n); s.out(lmax, dmax, n); s.out(n_steps, dns, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_define_dark_zone(mcb_handle* h, int32_t lambda, float tau_max, const double* r_grid, const double* z_grid,
                                 int32_t n_regions, const int32_t* region_iRmin, const int32_t* region_iRmax, const double* dust_density_sum,
                                 int32_t* l_dark_zone, int32_t* ri_in_dark_zone, int32_t* ri_out_dark_zone, int32_t* zj_sup_dark_zone,
                                 int32_t* zj_inf_dark_zone, int32_t* l_is_dark_zone) {
  if (!h || !r_grid || !z_grid || !l_dark_zone || !ri_in_dark_zone || !ri_out_dark_zone || !zj_sup_dark_zone) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "define_dark_zone before upload_grid/opacity");
  if (h->gk == GK_VOR) return fail(h, MCB_ERR_UNSUPPORTED, "define_dark_zone: structured grids only (the reference has no dark zone on a Voronoi mesh)");
  const DevModel& m = h->m;
  if (lambda < 1 || lambda > m.n_lambda) return fail(h, MCB_ERR_BAD_ARG, "lambda out of range");
  if (n_regions < 0 || (n_regions > 0 && (!region_iRmin || !region_iRmax))) return fail(h, MCB_ERR_BAD_ARG, "regions");
  for (int q = 0; q < n_regions; ++q)
    if (region_iRmin[q] < 1 || region_iRmin[q] > m.n_rad || region_iRmax[q] < 1 || region_iRmax[q] > m.n_rad) return fail(h, MCB_ERR_BAD_ARG, "region radius index out of range");
  if (m.l3D && !zj_inf_dark_zone) return fail(h, MCB_ERR_BAD_ARG, "zj_inf_dark_zone is needed on a 3D grid");
  if (!m.r_lim) return fail(h, MCB_ERR_STATE, "define_dark_zone: r_lim was not uploaded (mcb_grid::r_lim)");
  CK(cudaSetDevice(h->device));
  const int n_az = m.n_az > 0 ? m.n_az : 1;
  const int64_t nc = m.n_cells, nzs = (int64_t)m.n_rad * n_az;
  // cos / sin of the `real` angles (optical_depth.f90:1532,1557) with the host's single-precision libm
  std::vector<float> cs_ang(22), cs_phi(2 * (size_t)n_az);
  for (int n = 1; n <= 11; ++n) {
    const float angle = (float)(MCB_PI * (double)(float)n / (double)(float)12);
    cs_ang[2 * (n - 1)] = cosf(angle); cs_ang[2 * (n - 1) + 1] = sinf(angle);
  }
  for (int pk = 1; pk <= n_az; ++pk) {
    const float phi = (float)(2.0 * MCB_PI * (double)((float)pk - 0.5f) / (double)(float)n_az);
    cs_phi[2 * (pk - 1)] = cosf(phi); cs_phi[2 * (pk - 1) + 1] = sinf(phi);
  }
  std::vector<int32_t> zero_inf;
  if (!zj_inf_dark_zone) zero_inf.assign((size_t)nzs, 0);
  Scratch s{h};
  const double *drg = s.in(r_grid, nc), *dzg = s.in(z_grid, nc);
  const double* dds = dust_density_sum ? s.in(dust_density_sum, nc) : nullptr;
  const float *dang = s.in(cs_ang.data(), 22), *dphi = s.in(cs_phi.data(), 2 * (int64_t)n_az);
  const int *dmin = n_regions ? s.in(region_iRmin, n_regions) : nullptr, *dmax = n_regions ? s.in(region_iRmax, n_regions) : nullptr;
  int *ddark = s.in<int>(nullptr, nc), *drin = s.in<int>(nullptr, n_az), *drout = s.in<int>(nullptr, n_az);
  int *dzs = s.in(zj_sup_dark_zone, nzs), *dzi = s.in(zj_inf_dark_zone ? zj_inf_dark_zone : zero_inf.data(), nzs);
  int* dflag = s.in<int>(nullptr, 1);
  if (!drg || !dzg || !dang || !dphi || !ddark || !drin || !drout || !dzs || !dzi || !dflag) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
#define DDZ(GEOM) define_dark_zone_kernel<GEOM><<<1, 1024, 0, h->stream>>>(h->m, lambda, tau_max, drg, dzg, dang, dphi, dds, n_regions, dmin, dmax, ddark, drin, drout, dzs, dzi, dflag)
  switch (h->gk) {
    case GK_CYL2D: DDZ(GeomCyl<false>); break;
    case GK_CYL3D: DDZ(GeomCyl<true>); break;
    case GK_SPH2D: DDZ(GeomSph<false>); break;
    case GK_SPH3D: DDZ(GeomSph<true>); break;
    default: break;
  }
#undef DDZ
  CK(cudaGetLastError());
  s.out(l_dark_zone, ddark, nc); s.out(ri_in_dark_zone, drin, n_az); s.out(ri_out_dark_zone, drout, n_az);
  s.out(zj_sup_dark_zone, dzs, nzs);
  if (zj_inf_dark_zone) s.out(zj_inf_dark_zone, dzi, nzs);
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, dflag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (l_is_dark_zone) *l_is_dark_zone = flag;
  return mcfost_b200_upload_dark_zone(h, l_dark_zone);      // the photon loop of this handle uses the new dark zone from now on
}

int mcfost_b200_init_reemission(mcb_handle* h, const double* tab_lambda, const double* tab_delta_lambda,
                                double* log_Qcool_minus_extra_heating, double* kdB_dT_CDF) {
  if (!h || !tab_lambda || !tab_delta_lambda) return MCB_ERR_BAD_ARG;
  if (!h->has_op) return fail(h, MCB_ERR_STATE, "init_reemission before upload_opacity (kappa_abs_LTE, tab_Temp)");
  DevModel& m = h->m;
  if (!m.kappa_abs || !m.tab_Temp) return fail(h, MCB_ERR_STATE, "init_reemission: kappa_abs_LTE / tab_Temp missing");
  CK(cudaSetDevice(h->device));
  const int nl = m.n_lambda, nT = m.n_T, pnc = m.p_n_cells;
  Scratch s{h};
  const double *dl = s.in(tab_lambda, nl), *dd = s.in(tab_delta_lambda, nl);
  double *dB_ = s.in<double>(nullptr, (int64_t)nl * nT), *ddB = s.in<double>(nullptr, (int64_t)nl * nT);
  if (!dl || !dd || !dB_ || !ddB) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  double *dlogQ = nullptr, *dcdf = nullptr;
  int rc;
  if ((rc = reserve(h, "logQ", (size_t)nT * pnc, &dlogQ))) return rc;
  if ((rc = reserve(h, "kdB", (size_t)nl * nT * pnc, &dcdf))) return rc;
  planck_tables_kernel<<<(nl * nT + 127) / 128, 128, 0, h->stream>>>(nl, nT, dl, dd, m.tab_Temp, dB_, ddB);
  CK(cudaGetLastError());
  const int64_t n = (int64_t)nT * pnc;
  init_reemission_rows_kernel<false><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(nl, nT, pnc, m.kappa_abs, pnc, nullptr, 0, 0, dB_, ddB, dlogQ, nullptr, dcdf);
  CK(cudaGetLastError());
  m.logQ = dlogQ; m.kdB = dcdf;      // the thermal tables of this handle from now on (as upload_opacity would have set them)
  h->mrw_ready = false;
  if (log_Qcool_minus_extra_heating) CK(cudaMemcpyAsync(log_Qcool_minus_extra_heating, dlogQ, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (kdB_dT_CDF) CK(cudaMemcpyAsync(kdB_dT_CDF, dcdf, (size_t)n * nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_init_reemission_grains(mcb_handle* h, const double* tab_lambda, const double* tab_delta_lambda, const float* C_abs_norm,
                                       int32_t n_grains_tot, int32_t k_start, int32_t k_end, double* log_E_em_1grain, double* E_em_1grain,
                                       double* kdB_dT_1grain_CDF) {
  if (!h || !tab_lambda || !tab_delta_lambda || !C_abs_norm || !log_E_em_1grain || !kdB_dT_1grain_CDF) return MCB_ERR_BAD_ARG;
  if (!h->has_op) return fail(h, MCB_ERR_STATE, "init_reemission_grains before upload_opacity (n_lambda, tab_Temp)");
  const DevModel& m = h->m;
  if (k_start < 1 || k_end < k_start || k_end > n_grains_tot) return fail(h, MCB_ERR_BAD_ARG, "grain range");
  CK(cudaSetDevice(h->device));
  const int nl = m.n_lambda, nT = m.n_T, nk = k_end - k_start + 1;
  Scratch s{h};
  const double *dl = s.in(tab_lambda, nl), *dd = s.in(tab_delta_lambda, nl);
  const float* dca = s.in(C_abs_norm, (int64_t)n_grains_tot * nl);
  double *dB_ = s.in<double>(nullptr, (int64_t)nl * nT), *ddB = s.in<double>(nullptr, (int64_t)nl * nT);
  const int64_t n = (int64_t)nT * nk;
  double *dlogE = s.in<double>(nullptr, n), *dE = s.in<double>(nullptr, n), *dcdf = s.in<double>(nullptr, n * nl);
  if (!dl || !dd || !dca || !dB_ || !ddB || !dlogE || !dE || !dcdf) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  planck_tables_kernel<<<(nl * nT + 127) / 128, 128, 0, h->stream>>>(nl, nT, dl, dd, m.tab_Temp, dB_, ddB);
  CK(cudaGetLastError());
  init_reemission_rows_kernel<true><<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(nl, nT, nk, nullptr, 0, dca, n_grains_tot, k_start - 1, dB_, ddB, dlogE, dE, dcdf);
  CK(cudaGetLastError());
  s.out(log_E_em_1grain, dlogE, n); s.out(E_em_1grain, dE, n); s.out(kdB_dT_1grain_CDF, dcdf, n * nl);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_init_dust_source_fct1(mcb_handle* h, int32_t lambda, int32_t iRT, double photon_energy, const double* J_th, double* eps_dust1) {
  if (!h || !J_th) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "init_dust_source_fct1 before upload_grid/opacity");
  if (h->gk == GK_VOR) return fail(h, MCB_ERR_UNSUPPORTED, "init_dust_source_fct1: structured grids (the rt1 tally is tabulated on them)");
  DevModel& m = h->m;
  if (!h->n_xI || !h->rt1_n_rt) return fail(h, MCB_ERR_STATE, "init_dust_source_fct1: no xI_scatt tally on the device (run a step with lscatt_ray_tracing1 first)");
  if (lambda < 1 || lambda > m.n_lambda) return fail(h, MCB_ERR_BAD_ARG, "lambda out of range");
  if (iRT < 1 || iRT > h->rt1_n_rt) return fail(h, MCB_ERR_BAD_ARG, "iRT out of range (RT2d_to_RT1d(ibin, iaz), 1-based)");
  CK(cudaSetDevice(h->device));
  const int ntf = h->n_type_flux, n_stokes = h->rt1_pola ? 4 : 1;
  const int n_az_rt = m.l3D ? 1 : N_AZ_RT, n_theta_rt = m.l3D ? 1 : 2;
  const int64_t ne = (int64_t)N_AZ_RT * 2 * ntf * m.n_cells;
  double* deps = nullptr;
  int rc;
  if ((rc = reserve(h, "eps_dust1", (size_t)ne, &deps))) return rc;
  Scratch s{h};
  const double* dJ = s.in(J_th, m.n_cells);
  if (!dJ) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const int64_t nq = (int64_t)N_AZ_RT * 2 * m.n_cells;
  init_dust_source_fct1_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, h->stream>>>(m, lambda, iRT, h->rt1_n_rt, photon_energy, dJ, m.xI, n_az_rt,
                                                                                  n_theta_rt, ntf, n_stokes, h->rt1_pola, h->rt1_contrib, deps);
  CK(cudaGetLastError());
  if (eps_dust1) CK(cudaMemcpyAsync(eps_dust1, deps, (size_t)ne * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->eps_ntf = ntf; h->eps_lambda = lambda;
  return MCB_OK;
}

int mcfost_b200_integ_ray_dust(mcb_handle* h, int32_t lambda, int64_t n, const double* x, const double* y, const double* z, const double* u,
                               const double* v, const double* w, const int32_t* icell, float tau_dark_zone_obs, double* I) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "integ_ray_dust before upload_grid/opacity");
  if (h->gk == GK_VOR) return fail(h, MCB_ERR_UNSUPPORTED, "integ_ray_dust: structured grids");
  if (!h->eps_ntf || !h->bufs.count("eps_dust1")) return fail(h, MCB_ERR_STATE, "integ_ray_dust before init_dust_source_fct1");
  if (lambda != h->eps_lambda) return fail(h, MCB_ERR_BAD_ARG, "integ_ray_dust: the source function on the device was built for another wavelength");
  if (n == 0) return MCB_OK;
  if (!x || !y || !z || !u || !v || !w || !icell || !I) return MCB_ERR_BAD_ARG;
  CK(cudaSetDevice(h->device));
  const int ntf = h->eps_ntf;
  Scratch s{h};
  const double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n), *du = s.in(u, n), *dv = s.in(v, n), *dw = s.in(w, n);
  const int* dic = s.in(icell, n);
  double* dout = s.in<double>(nullptr, n * ntf);
  if (!dx || !dy || !dz || !du || !dv || !dw || !dic || !dout) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const double* deps = (const double*)h->bufs["eps_dust1"];
  const int n_az_rt = h->m.l3D ? 1 : N_AZ_RT;
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(integ_ray_dust_kernel, h->m, n, lambda, dx, dy, dz, du, dv, dw, dic, tau_dark_zone_obs, deps, n_az_rt, ntf, dout);
  CK(cudaGetLastError());
  s.out(I, dout, n * ntf);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_repartition_energie(mcb_handle* h, int32_t lambda_first, int32_t lambda_last, const float* Tdust, const double* tab_lambda,
                                    const double* E_stars, const double* E_ISM, const double* weight_proba_emission, double* E_disk,
                                    double* frac_E_stars, double* frac_E_disk, double* weight_norm, double* prob_E_cell) {
  if (!h || !Tdust || !tab_lambda || !E_stars || !E_disk || !frac_E_stars || !frac_E_disk) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "repartition_energie before upload_grid/opacity");
  DevModel& m = h->m;
  if (lambda_first < 1 || lambda_last > m.n_lambda || lambda_last < lambda_first) return fail(h, MCB_ERR_BAD_ARG, "wavelength range");
  if (!m.volume) return fail(h, MCB_ERR_STATE, "repartition_energie: cell volumes were not uploaded");
  CK(cudaSetDevice(h->device));
  const int nc = m.n_cells, nl = m.n_lambda;
  int rc;
  double *dprob = nullptr, *dfs = nullptr, *dfd = nullptr;
  // the emission tables of the handle (what upload_emission fills); created on first use
  const bool had = h->bufs.count("prob_E_cell") && h->buf_bytes["prob_E_cell"] == (size_t)(nc + 1) * nl * sizeof(double);
  if ((rc = reserve(h, "prob_E_cell", (size_t)(nc + 1) * nl, &dprob))) return rc;
  if ((rc = reserve(h, "frac_star", (size_t)nl, &dfs))) return rc;
  if ((rc = reserve(h, "frac_disk", (size_t)nl, &dfd))) return rc;
  if (!had) CK(cudaMemsetAsync(dprob, 0, (size_t)(nc + 1) * nl * sizeof(double), h->stream));
  Scratch s{h};
  const float* dT = s.in(Tdust, nc);
  const double* dw = weight_proba_emission ? s.in(weight_proba_emission, nc) : nullptr;
  double *dE = s.in<double>(nullptr, nc), *dEc = s.in<double>(nullptr, nc), *dcum = s.in<double>(nullptr, nc), *dsum = s.in<double>(nullptr, 2);
  if (!dT || !dE || !dEc || !dcum || !dsum || (weight_proba_emission && !dw)) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  size_t tmp_scan = 0, tmp_red = 0, tmp_max = 0;
  cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, dEc, dcum, nc, h->stream);
  cub::DeviceScan::InclusiveScan(nullptr, tmp_max, dcum, dE, cub::Max(), nc, h->stream);
  if (tmp_max > tmp_scan) tmp_scan = tmp_max;
  cub::DeviceReduce::Sum(nullptr, tmp_red, dE, dsum, nc, h->stream);
  void* dtmp = nullptr;
  CK(cudaMalloc(&dtmp, tmp_scan > tmp_red ? tmp_scan : tmp_red));
  s.d.push_back(dtmp);
  const unsigned nb = (unsigned)((nc + 127) / 128), nb1 = (unsigned)((nc + 1 + 127) / 128);
  std::vector<double> fs((size_t)nl), fd((size_t)nl);
  CK(cudaMemcpyAsync(fs.data(), dfs, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));      // keep the wavelengths outside the range
  CK(cudaMemcpyAsync(fd.data(), dfd, (size_t)nl * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int l = lambda_first; l <= lambda_last; ++l) {
    const double wl = tab_lambda[l - 1] * (double)1.e-6f;
    emission_cells_kernel<<<nb, 128, 0, h->stream>>>(m, l, wl, dT, dw, dE, dEc);
    CK(cudaGetLastError());
    size_t t1 = tmp_scan, t2 = tmp_red;
    cub::DeviceScan::InclusiveSum(dtmp, t1, dEc, dcum, nc, h->stream);
    cub::DeviceReduce::Sum(dtmp, t2, dE, dsum, nc, h->stream);
    // a parallel prefix sum is not monotone to the last bit (each element has its own association order): a running
    // maximum makes the distribution function non-decreasing again (dE is free once its sum is taken)
    size_t t3 = tmp_scan;
    cub::DeviceScan::InclusiveScan(dtmp, t3, dcum, dE, cub::Max(), nc, h->stream);
    normalise_cdf_kernel<<<nb1, 128, 0, h->stream>>>(dE, nc, dprob + (size_t)(nc + 1) * (l - 1));
    CK(cudaGetLastError());
    double sums[2] = {0.0, 0.0};
    CK(cudaMemcpyAsync(&sums[0], dsum, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&sums[1], dE + (nc - 1), sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double Ed = sums[0], Es = E_stars[l - 1], Ei = E_ISM ? E_ISM[l - 1] : 0.0;
    if (Es + Ed + Ei < MCB_TINY_DP) return fail(h, MCB_ERR_BAD_ARG, "repartition_energie: no energy at this wavelength (the reference exits, thermal_emission.f90:1900)");
    E_disk[l - 1] = Ed;
    frac_E_stars[l - 1] = fs[l - 1] = Es / (Es + Ed + Ei);
    frac_E_disk[l - 1] = fd[l - 1] = (Es + Ed) / (Es + Ed + Ei);
    if (weight_norm) weight_norm[l - 1] = Ed > 0.0 ? sums[1] / Ed : 0.0;
    if (prob_E_cell) CK(cudaMemcpyAsync(prob_E_cell + (size_t)(nc + 1) * (l - 1), dprob + (size_t)(nc + 1) * (l - 1), (size_t)(nc + 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CK(cudaMemcpyAsync(dfs, fs.data(), (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemcpyAsync(dfd, fd.data(), (size_t)nl * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  m.prob_E_cell = dprob; m.frac_star = dfs; m.frac_disk = dfd;
  h->em_on_device = true;
  return MCB_OK;
}

int mcfost_b200_compute_column(mcb_handle* h, int32_t lambda, const double* factor, const double* centre_x, const double* centre_y,
                               const double* centre_z, float* column) {
  if (!h || !centre_x || !centre_y || !centre_z || !column) return MCB_ERR_BAD_ARG;
  if (!h->has_grid) return fail(h, MCB_ERR_STATE, "compute_column before upload_grid");
  if (!factor) {
    if (!h->has_op) return fail(h, MCB_ERR_STATE, "compute_column (optical depth) before upload_opacity");
    if (lambda < 1 || lambda > h->m.n_lambda) return fail(h, MCB_ERR_BAD_ARG, "lambda out of range");
  }
  const int64_t nc = h->m.n_cells, n = 4 * nc;
  if (nc == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  const double *dx = s.in(centre_x, nc), *dy = s.in(centre_y, nc), *dz = s.in(centre_z, nc);
  const double* df = factor ? s.in(factor, nc) : nullptr;
  float* dcol = s.in<float>(nullptr, n);
  if (!dx || !dy || !dz || !dcol || (factor && !df)) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(compute_column_kernel, h->m, lambda, df, dx, dy, dz, dcol);
  CK(cudaGetLastError());
  s.out(column, dcol, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

int mcfost_b200_physical_length(mcb_handle* h, int64_t n, int32_t lambda, double* x, double* y, double* z, double* u, double* v, double* w,
                                int32_t* icell, const float* tau, float* ltot, int32_t* flag_sortie, int32_t* lpacket_alive) {
  if (!h || n < 0) return MCB_ERR_BAD_ARG;
  if (!h->has_grid || !h->has_op) return fail(h, MCB_ERR_STATE, "physical_length before upload_grid/opacity");
  if (lambda < 1 || lambda > h->m.n_lambda) return fail(h, MCB_ERR_BAD_ARG, "lambda out of range");
  if (n == 0) return MCB_OK;
  CK(cudaSetDevice(h->device));
  Scratch s{h};
  double *dx = s.in(x, n), *dy = s.in(y, n), *dz = s.in(z, n), *du = s.in(u, n), *dv = s.in(v, n), *dw = s.in(w, n);
  int* dic = s.in(icell, n);
  const float* dtau = s.in(tau, n);
  float* dl = s.in<float>(nullptr, n);
  int *dfs = s.in<int>(nullptr, n), *dal = s.in<int>(nullptr, n);
  if (!dal) return fail(h, MCB_ERR_CUDA, "scratch allocation failed");
  const unsigned nb = (unsigned)((n + 127) / 128);
  DISPATCH(physical_length_kernel, h->m, n, lambda, dx, dy, dz, du, dv, dw, dic, dtau, dl, dfs, dal);
  CK(cudaGetLastError());
  s.out(x, dx, n); s.out(y, dy, n); s.out(z, dz, n); s.out(u, du, n); s.out(v, dv, n); s.out(w, dw, n);
  s.out(icell, dic, n); s.out(ltot, dl, n); s.out(flag_sortie, dfs, n); s.out(lpacket_alive, dal, n);
  CK(cudaStreamSynchronize(h->stream));
  return MCB_OK;
}

}  // extern "C"
