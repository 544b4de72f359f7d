// Device-side view of everything uploaded through the C ABI (include/mcfost_b200.h).
// One DevModel per handle, passed to kernels by value as a __grid_constant__
// parameter (constant bank, broadcast to every warp).
#pragma once
#include <cstdint>
#include <cfloat>

namespace mcb {

constexpr int MAX_STARS = 15;      // the index of the star a flight points at has 4 bits in the packed packet state (0 = none)
constexpr int NANG = 180;      // nang_scatt
constexpr int N_AZ_RT = 45;    // n_az_rt
constexpr int MAX_RT = 16;     // max RT_n_incl * RT_n_az observer directions
constexpr int N_ZETA = 10000;  // MRW.f90:8

// numerical constants of the reference (constants.f90:8-14,151-159)
#define MCB_PI       3.141592653589793238462643383279502884197
#define MCB_TWO_PI   (2.0 * MCB_PI)
#define MCB_HALF_PI  (0.5 * MCB_PI)
#define MCB_TINY_REAL ((double)FLT_MIN)
#define MCB_HUGE_REAL ((double)FLT_MAX)
#define MCB_HUGE_DP   DBL_MAX
#define MCB_TINY_DP   DBL_MIN
#define MCB_GRID_PREC 1.0e-14            /* cylindrical_grid.f90:16 */
#define MCB_PREC_SPH  1.0e-7             /* spherical_grid.f90:19   */

enum GridKind { GK_CYL2D = 0, GK_CYL3D = 1, GK_SPH2D = 2, GK_SPH3D = 3, GK_VOR = 4 };

// tally block offsets inside the packed fp64 buffer
struct TallyLayout {
  int64_t xKJ, xJ, n_env, sed, stats, E_abs_nRE, total;   // offsets in doubles
  int64_t n_sed;                               // n_lambda*N_thet*N_phi
};

// Shared-memory staging of the small hot tables (offsets in 8-byte words from the
// start of dynamic shared memory; float tables are stored widened to their own
// word-aligned region).  Used by the photon-loop kernel when lvariable_dust is off
// (p_n_cells == 1) and the tables fit; the per-cell arrays (kappa_factor,
// l_dark_zone, volume, tallies) stay in global memory behind L1/L2.
struct SmemLayout {
  int enabled;
  int r_lim_2, zmax, zl, tan_phi, tan_theta;          // geometry (zl = cell_height if z is regular else z_lim)
  int kappa, kappa_abs, albedo, gfac;                 // (n_lambda); albedo/gfac are float regions
  int logQ, kdB, cos_tab, prob_s11;                   // thermal / scattering (prob_s11: float region, one p_lambda slice)
  int spec_cumul, frac_star, frac_disk;
  int total_words;
};

// per-grain tables (mcb_grains): scattering method 1, nLTE / qRE re-emission
struct DevGrains {
  int n_grains_tot, n_dens;
  int LTE_s, LTE_e, nLTE_s, nLTE_e, nRE_s, nRE_e;     // 1-based inclusive grain ranges (grains.f90:36)
  const int *zone;                  // (n_grains_tot)
  const double *n_grains;           // (n_grains_tot)
  const double *dd;                 // dust_density_o_n_grains (n_dens, n_cells)
  const float *C_abs, *C_abs_norm, *C_sca, *tab_g;    // (n_grains_tot, n_lambda)
  const float *prob_s11;            // (n_lambda, n_grains_tot, 0:180)
  const float *s11, *s12, *s22, *s33, *s34, *s44;     // (0:180, n_grains_tot, n_lambda)
  const double *ksca_CDF;           // (0:n_grains_tot, p_n_cells, n_lambda)
  const double *kappa_abs_nLTE;     // (p_n_cells, n_lambda)
  const double *kabs_nLTE_CDF;      // (nLTE_s-1:nLTE_e, n_cells, n_lambda)
  const double *logE, *kdB;         // nLTE grains: (k, n_T), (n_lambda, k, n_T)
  const double *kappa_abs_RE, *proba_abs_RE, *P_LTE, *P_LTE_p_nLTE;   // (n_cells, n_lambda)
  const double *logE_nRE, *kdB_nRE; // nRE grains
  const int *l_RE;                  // (nRE grains, n_cells)
  const double *J0;                 // (n_cells, n_lambda)
  const double *kdB_LTE;            // LTE grains, low-memory emission: (n_lambda, k, n_T)
  int *xT_1g, *xT_1g_nRE;           // tallies: xT_ech_1grain (nLTE grains, n_cells), xT_ech_1grain_nRE
};

struct DevModel {
  // ---- grid (cylindrical_grid.f90:20-41) --------------------------------
  int kind, l3D, n_rad, nz, n_az, n_cells, nj;   // nj = rows per azimuth (nz or 2 nz)
  int z_regular;                                 // z_lim(i,j) == (j-1)*z_lim(i,2) bit-for-bit (default grid, :459-465)
  double Rmax2, zmaxmax;
  const double *r_lim_2, *r_lim_3, *z_lim, *zmax, *cell_height, *tan_theta_lim, *theta_lim, *tan_phi_lim, *volume;
  // walls as read by distance_to_closest_wall_* (modified random walk): r_lim(0:n_rad), w_lim(0:nz) = sin(theta_lim),
  // cos(theta_lim(0:nz)) (host libm), sin / cos_phi_lim(n_az)
  const double *r_lim, *w_lim, *cos_theta_lim, *sin_phi_lim, *cos_phi_lim;
  const double *kappa_factor;       // (n_cells)
  const double *kf_dark;            // (n_cells) kappa_factor with the sign bit set where l_dark_zone (photon-loop kernel only)
  const uint8_t *dark;              // (n_cells) l_dark_zone
  // ---- Voronoi (Voronoi.f90:23-66) --------------------------------------
  const double *vor_xyz;            // (3, n_cells) fp64
  const float4 *vor_xyz32;          // fp32 copy (Voronoi_xyz, :61), one float4 per cell
  const double *vor_h;
  const int *vor_first, *vor_last, *neigh;
  const uint8_t *vor_flags;         // bit0 was_cut, bit1 is_star, bit2 is_star_neighbour
  float wall[6][4];
  double cut_o_h;
  // uniform grid over the seeds' bounding box for point location (the reference uses a kd-tree, Voronoi.f90:1625-1645):
  // vg_start(0:nx*ny*nz), vg_items = seed ids (1-based, ascending inside a grid cell)
  const int *vg_start, *vg_items;
  int vg_n[3];
  double vg_lo[3], vg_inv[3], vg_step[3];
  // ---- stars ------------------------------------------------------------
  int n_stars;              // (the star tables are at the end of the struct: measured, 15 stars in the middle of it cost the packet-per-lane kernel 1.2 %)
  // ---- opacity ----------------------------------------------------------
  int n_lambda, p_n_cells, p_n_lambda_pos, n_T;
  const double *kappa, *kappa_abs;          // (p_n_cells, n_lambda)
  const float *albedo, *gfac;               // (p_n_cells, n_lambda)
  const float *prob_s11, *s11, *s12, *s22, *s33, *s34, *s44;   // (0:180, p_n_cells, p_n_lambda_pos)
  const double *logQ;                       // (n_T, p_n_cells)
  const double *kdB;                        // (n_lambda, n_T, p_n_cells)
  const double *cos_tab;                    // cos(k*pi/180), k = 0..180 (host libm)
  float T_min;
  const float *tab_Temp;                    // (n_T), read by the per-grain re-emission branches only
  // ---- emission ---------------------------------------------------------
  const double *spec_cumul, *frac_star, *frac_disk, *prob_E_cell;
  const float *CDF_E_star;
  double L_packet_th, E_paquet, R_ISM, cISM[3];
  const double *correct_E;          // (n_cells) correct_E_emission (lweight_emission) or null
  const double *tab_lambda;         // (n_lambda) micron (hot spot only)
  // ---- tallies ----------------------------------------------------------
  double *tally;            // packed fp64 block
  TallyLayout lay;
  int *xT_ech;              // (n_cells)
  float *xI;                // xI_scatt
  float *I_spec, *I_spec_star;   // rt2 accumulators (dust_ray_tracing.f90:44-45)
  double *quv;              // Stokes Q,U,V of the packets in flight: (n_blocks, 3, NP), only with lsepar_pola
  unsigned long long *work; // [0] = next work item; [2+2c], [3+2c] = sent / received of local chunk c
  double *pos0;             // (n_blocks, 4, NP): start point x,y,z and cell index of the flight in progress (photon maps / lorigine only)
  double *xN;               // xN_abs: (n_cells) thermal / (n_cells, n_lambda) otherwise, or null
  double *smap, *star_origin, *disk_origin;   // Monte Carlo photon maps of the call's wavelength, packet-origin tallies (output.f90:26-37)
  double *park;             // parked stragglers: (PARK_REC doubles) x capacity, see transport.cuh
  // modified random walk (MRW.f90): zeta(1:n_zeta), mean opacities A, B, C (n_T, p_n_cells), flight-start cell ids
  const double *zeta, *mrw_A, *mrw_B, *mrw_C;
  float *mrw_lR;            // (2, n_cells): mean free path of the last walk evaluated in the cell (0 = none yet) and the temperature index it was evaluated at, see mrw_worth_trying
  SmemLayout sm;
  DevGrains gr;
  double star[MAX_STARS][4];
  int star_icell[MAX_STARS], star_out[MAX_STARS];
};

// run parameters broadcast to the kernel
struct DevRun {
  int lambda_in, p_lambda_in, n_photons2, nnfot1_start, n_photons_loop;
  float n_phot_lim;
  int letape_th, lmono, lsepar_pola, lsepar_contrib, lmethod_aniso1, lisotropic;
  int l_sym_centrale, l_sym_axiale, rt1, lxJ;
  int rt2, lmono0, n_theta_I, n_phi_I;
  int lscattering_method1, low_mem_scattering;          // dust_transfer.f90:1291, dust_prop.f90:1305
  int lonly_LTE, lonly_nLTE, lRE_nLTE, lnRE, low_mem_nLTE;   // grain heating regimes (dust_transfer.f90:1353-1395)
  int N_thet, N_phi, capt_sup, n_type_flux, n_stokes;
  int n_rt, RT_n_incl, RT_n_az;
  double rt_u[MAX_RT], rt_v[MAX_RT], rt_w[MAX_RT];
  unsigned long long seed;
  unsigned int call_index;
  int rank, n_ranks, n_local_chunks;
  int count_sent;                       // 1: chunk ends after n_photons2 packets SENT (thermal / image), 0: RECEIVED (SED)
  unsigned long long sent_lim;          // ceil(n_phot_lim) (saturated)
  unsigned long long n_per_chunk;       // count_sent: min(n_photons2, sent_lim)
  unsigned long long n_packets_total;   // count_sent: n_local_chunks * n_per_chunk
  double nb_proc_equiv;                 // n_ranks: scales the local tally in Temp_LTE
  // emission extras (dust_transfer.f90:1090-1142); spot direction and opening precomputed on the host
  int low_mem_th, lweight_emission, lspot, lxN;
  float x_spot, y_spot, z_spot, cos_thet_spot, T_spot;
  double star1_T;
  // capteur extras (output.f90:303-357,396-570)
  int capt_full;                        // 1: lorigine / lonly_capt_interet / photon maps are on -> capteur_full
  int mc_maps, lorigine, capt_interet, lonly_capt_interet, capt_inf, npix_x, npix_y, l_sym_ima;
  double zoom, map_size, cos_disk, sin_disk;
  int patience;                         // polls (250 ns each) a warp waits for a full 32-packet chunk before it takes a partial one
  int patience_dry, drain_live_dry;     // the same two thresholds once the packet counter ran dry
  int min_chunk;                        // a chunk of at least this many packets is taken without waiting (32: only full chunks)
  int park_live;                        // hand over when at most this many packets of a block are in flight (<= PARK_LIVE)
  int park_enable;                      // hand stragglers over to a second small launch (count_sent modes only)
  int debug_abort_dry;                  // profiling aid, development builds (-DMCB_DEV) only: stop when the packet counter runs dry
  int lMRW;                             // modified random walk in the thermal step
  int lism;                             // ISM side loop (dust_transfer.f90:941-985): every packet from emit_packet_ISM, chunks count the packets that enter the model
  double gamma_MRW;
  unsigned inflight_floor;              // packets in flight per block: never capped below this,
  float inflight_frac_per_block;        // else max_inflight_fraction * packets sent so far / blocks
};

}  // namespace mcb
