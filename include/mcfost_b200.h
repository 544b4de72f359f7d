/*
 * mcfost_b200.h -- C ABI of the B200-native replacement for MCFOST's Monte Carlo
 * photon-packet loop.
 *
 * What it replaces (reference = cpinte/mcfost 4.1.13, paths relative to src/):
 *   the body of  subroutine mc_photon_loop(lambda_in, p_lambda_in, n_photons2,
 *   n_phot_lim, nnfot1_start, laffichage)            dust_transfer.f90:439-572
 *   and everything it calls (emit_packet :1047, propagate_packet :1155,
 *   physical_length optical_depth.f90:21, cross_cell / index_cell / move_to_grid
 *   of cylindrical_grid.f90, spherical_grid.f90, Voronoi.f90, save_radiation_field
 *   radiation_field.f90:31, im_reemission_LTE thermal_emission.f90:710, the
 *   scattering samplers scattering.f90:1187-1475 and capteur output.f90:294).
 *
 * Conventions follow the reference's own C boundary (voro_C, Voronoi.f90:70-96
 * <-> voro++_wrapper.cpp:41-44): plain pointers and sizes, the CALLER allocates
 * and owns every host buffer, the callee never frees caller memory, and every
 * entry point returns an int error code (0 = OK; non-zero -> the Fortran shim
 * calls error() -> exit(1), messages.f90:27-46).
 *
 * Layout: all arrays are Fortran column-major exactly as the reference's module
 * variables hold them, cell ids are the reference's 1-based ids (virtual cells
 * n_cells+1..ntot2 for cylindrical/spherical grids, negative ids for Voronoi
 * walls), wavelength / temperature / grain indices are 1-based.  Fortran
 * `logical` arrays cross as int32 (0 / non-zero).
 *
 * The same structs are consumed by the CPU oracle (oracle/, test infrastructure
 * only) so that both sides are fed byte-identical inputs.
 */
#ifndef MCFOST_B200_H
#define MCFOST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------- */
#define MCB_OK                0
#define MCB_ERR_NO_DEVICE     1   /* no CUDA device / driver: there is NO CPU fallback */
#define MCB_ERR_BAD_ARG       2
#define MCB_ERR_CELL_MAP      3   /* caller's cell_map_{i,j,k} disagree with the analytic numbering */
#define MCB_ERR_CUDA          4
#define MCB_ERR_UNSUPPORTED   5   /* a mode flag the library does not implement (yet): fail loudly */
#define MCB_ERR_STATE         6   /* run called before the uploads it needs */

/* ---- grid kinds (grid.f90:273-367 binds the procedure pointers by kind) */
#define MCB_GRID_CYL      1   /* lcylindrical, cylindrical_grid.f90 */
#define MCB_GRID_SPH      2   /* lspherical,   spherical_grid.f90   */
#define MCB_GRID_VORONOI  3   /* lVoronoi,     Voronoi.f90          */

#define MCB_NANG_SCATT  180   /* nang_scatt, parameters.f90 */
#define MCB_N_AZ_RT      45   /* n_az_rt, dust_ray_tracing.f90:33 */

typedef struct mcb_handle mcb_handle;   /* opaque; owns all device memory */

/* ------------------------------------------------------------------------
 * Grid + stars.  Mirrors the module variables of cylindrical_grid.f90:20-41,
 * Voronoi.f90:23-66 and stars (star(:)%x,y,z,r,icell,out_model).
 * --------------------------------------------------------------------- */
typedef struct mcb_grid {
  int32_t kind;            /* MCB_GRID_* */
  int32_t l3D;             /* l3D: j runs -nz..-1,1..nz (grid.f90:279-283) */
  int32_t n_rad, nz, n_az;
  int32_t n_cells;         /* real cells */
  double  Rmax2;           /* Rmax**2, cylindrical_grid.f90:254 */
  double  zmaxmax;         /* maxval(zmax), :494 (cyl only) */

  /* cylindrical / spherical tables */
  const double *r_lim;          /* (0:n_rad)            */
  const double *r_lim_2;        /* (0:n_rad)            */
  const double *r_lim_3;        /* (0:n_rad)  (sph)     */
  const double *z_lim;          /* (n_rad, nz+2) column-major (cyl) */
  const double *zmax;           /* (n_rad)              (cyl) */
  const double *tan_theta_lim;  /* (0:nz)               (sph) */
  const double *theta_lim;      /* (0:nz)               (sph) */
  const double *tan_phi_lim;    /* (n_az)               (3D)  */
  const double *volume;         /* (n_cells)  AU^3      */

  /* optional: the reference's own numbering, used ONLY to verify that the
   * library's analytic numbering (build_cylindrical_cell_mapping,
   * cylindrical_grid.f90:45-179) matches; may be NULL. length n_cells_tot */
  int32_t        n_cells_tot;   /* ntot2 (real + virtual); 0 if maps are NULL */
  const int32_t *cell_map_i, *cell_map_j, *cell_map_k;

  /* Voronoi (kind==3): Voronoi(:)%xyz,h,first/last_neighbour,flags */
  const double  *vor_xyz;          /* (3, n_cells) */
  const double  *vor_h;            /* (n_cells)    */
  const int32_t *vor_first, *vor_last;   /* (n_cells) 1-based into neighbours_list */
  const int32_t *vor_was_cut, *vor_is_star, *vor_is_star_neighbour; /* (n_cells) */
  const int32_t *neighbours_list;  /* 1-based cell ids; <0 = -wall id */
  int64_t        n_neighbours_tot;
  float          wall_x[6][4];     /* wall(i)%x1,x2,x3,x4 (real) */
  double         cutting_distance_o_h;   /* PS%cutting_distance_o_h */

  /* stars */
  int32_t        n_stars;
  const double  *star_xyzr;        /* (4, n_stars): x,y,z,r  [AU] */
  const int32_t *star_icell;       /* (n_stars) star(:)%icell */
  const int32_t *star_out_model;   /* (n_stars) star(:)%out_model */

  /* wall tables read only by distance_to_closest_wall_* (modified random walk); may be NULL when lMRW is
   * never requested.  cylindrical_grid.f90:30-31,498-598 */
  const double *w_lim;             /* (0:nz)  sin(theta_lim)   (sph) */
  const double *sin_phi_lim, *cos_phi_lim;   /* (n_az)        (3D)  */
} mcb_grid;

/* ------------------------------------------------------------------------
 * Dust opacity / scattering / thermal tables.
 * p_n_cells = n_cells if lvariable_dust else 1 (grid.f90:292-296).
 * --------------------------------------------------------------------- */
typedef struct mcb_opacity {
  int32_t n_lambda;
  int32_t p_n_cells;
  int32_t p_n_lambda_pos;     /* n_lambda or 1 (scattering.f90:39-66) */
  int32_t n_T;

  const double *kappa;            /* (p_n_cells, n_lambda)  AU^-1  dust_prop.f90:17 */
  const double *kappa_abs_LTE;    /* (p_n_cells, n_lambda)  */
  const double *kappa_factor;     /* (n_cells)   dust_prop.f90:951-955 */
  const float  *tab_albedo_pos;   /* (p_n_cells, n_lambda)  grains.f90:62 */
  const float  *tab_g_pos;        /* (p_n_cells, n_lambda)  */

  /* method-2 scattering tables, (0:180, p_n_cells, p_n_lambda_pos), real */
  const float *prob_s11_pos;
  const float *tab_s11_pos;
  const float *tab_s12_o_s11_pos, *tab_s22_o_s11_pos, *tab_s33_o_s11_pos,
              *tab_s34_o_s11_pos, *tab_s44_o_s11_pos;   /* NULL unless lsepar_pola */

  /* Bjorkman & Wood tables (thermal_emission.f90:404-644) */
  const double *log_Qcool_minus_extra_heating;  /* (n_T, p_n_cells) */
  const double *kdB_dT_CDF;                     /* (n_lambda, n_T, p_n_cells) */
  const float  *tab_Temp;                       /* (n_T)  Temperature.f90:23-39 */
  float  T_min;                                 /* parameters: T_min */
} mcb_opacity;

/* ------------------------------------------------------------------------
 * Per-grain tables: scattering method 1 (dust_transfer.f90:1291-1317) and the
 * nLTE / qRE re-emission branches (dust_transfer.f90:1353-1395,
 * thermal_emission.f90:775-866,1441-1514,1953-2040).  Any pointer may be NULL
 * when the mode that reads it is off.  Grain indices are 1-based and global
 * (1..n_grains_tot); tables allocated on a sub-range of grains by the reference
 * (grain_RE_nLTE_start:grain_RE_nLTE_end, grain_nRE_start:grain_nRE_end) cross
 * with exactly that extent.
 * --------------------------------------------------------------------- */
typedef struct mcb_grains {
  int32_t n_grains_tot;
  int32_t n_dens;                  /* first extent of dust_density_o_n_grains: n_grains_tot if lvariable_dust else n_zones (mem.f90:35-39) */
  int32_t grain_RE_LTE_start, grain_RE_LTE_end;     /* grains.f90:36 */
  int32_t grain_RE_nLTE_start, grain_RE_nLTE_end;
  int32_t grain_nRE_start, grain_nRE_end;
  const int32_t *grain_zone;       /* (n_grains_tot) grain(k)%zone */
  const double  *n_grains;         /* (n_grains_tot) grains.f90:38 */
  const double  *dust_density_o_n_grains;  /* (n_dens, n_cells) density.f90:32 */
  const float   *C_abs, *C_abs_norm, *C_sca, *tab_g;   /* (n_grains_tot, n_lambda) grains.f90:54 */
  /* scattering method 1 */
  const float   *prob_s11;         /* (n_lambda, n_grains_tot, 0:180) mem.f90:108 */
  const float   *tab_s11, *tab_s12, *tab_s22, *tab_s33, *tab_s34, *tab_s44;  /* (0:180, n_grains_tot, n_lambda) mem.f90:84-104 */
  const double  *ksca_CDF;         /* (0:n_grains_tot, p_n_cells, n_lambda) mem.f90:251; NULL if low_mem_scattering */
  /* RE - nLTE grains */
  const double  *kappa_abs_nLTE;   /* (p_n_cells, n_lambda) dust_prop.f90:20 */
  const double  *kabs_nLTE_CDF;    /* (grain_RE_nLTE_start-1:grain_RE_nLTE_end, n_cells, n_lambda) thermal_emission.f90:147 */
  const double  *log_E_em_1grain;  /* (grain_RE_nLTE_start:grain_RE_nLTE_end, n_T) :159 */
  const double  *kdB_dT_1grain_nLTE_CDF;   /* (n_lambda, nLTE grains, n_T) :155 */
  /* nRE grains that reached quasi radiative equilibrium */
  const double  *kappa_abs_RE;     /* (n_cells, n_lambda) mem.f90:201 */
  const double  *proba_abs_RE, *Proba_abs_RE_LTE, *Proba_abs_RE_LTE_p_nLTE;   /* (n_cells, n_lambda) dust_prop.f90:21 */
  const double  *log_E_em_1grain_nRE;      /* (grain_nRE_start:grain_nRE_end, n_T) */
  const double  *kdB_dT_1grain_nRE_CDF;    /* (n_lambda, nRE grains, n_T) :190 */
  const int32_t *l_RE;             /* (grain_nRE_start:grain_nRE_end, n_cells) logical, :304 */
  const double  *J0;               /* (n_cells, n_lambda) radiation_field.f90:154 */
  /* LTE grains, low-memory emission (low_mem_th_emission, thermal_emission.f90:127-140,739-751) */
  const double  *kdB_dT_1grain_LTE_CDF;    /* (n_lambda, grain_RE_LTE_start:grain_RE_LTE_end, n_T) :137 */
} mcb_grains;

/* ------------------------------------------------------------------------
 * Emission tables (repartition_energie thermal_emission.f90:1771,
 * repartition_wl_em :315, stars.f90:495-605).  Re-uploaded whenever the
 * Fortran side recomputes them (each temperature iteration / wavelength).
 * --------------------------------------------------------------------- */
typedef struct mcb_emission {
  const double *spectre_emission_cumul;  /* (0:n_lambda)           */
  const double *frac_E_stars;            /* (n_lambda)             */
  const double *frac_E_disk;             /* (n_lambda)             */
  const double *prob_E_cell;             /* (0:n_cells, n_lambda)  */
  const float  *CDF_E_star;              /* (n_lambda, 0:n_stars)  */
  double  L_packet_th;                   /* thermal_emission.f90:355-356 */
  double  E_paquet;                      /* Stokes(1) at emission  */
  double  R_ISM;                         /* stars.f90:728-787      */
  double  centre_ISM[3];
  const double *correct_E_emission;      /* (n_cells) with lweight_emission (dust_transfer.f90:1140), else NULL */
} mcb_emission;

/* ------------------------------------------------------------------------
 * One mc_photon_loop call.  The six leading members are the subroutine's own
 * dummy arguments (dust_transfer.f90:439-454); the rest are the module-level
 * mode flags / scalars it reads.
 * --------------------------------------------------------------------- */
typedef struct mcb_run_params {
  int32_t lambda_in;        /* 1-based */
  int32_t p_lambda_in;      /* 1-based; frozen for the whole call (:490-502) */
  int32_t n_photons2;
  float   n_phot_lim;
  int32_t nnfot1_start;     /* 1-based first chunk */
  int32_t laffichage;       /* progress bar: ignored on device */

  int32_t n_photons_loop;   /* 128, read_param.f90:145 */
  /* mode flags (parameters.f90) */
  int32_t letape_th, lmono, lmono0;
  int32_t lscatt_ray_tracing1, lscatt_ray_tracing2;
  int32_t lsepar_pola, lsepar_contrib;
  int32_t lscattering_method1;   /* 1: per-grain scattering (needs mcfost_b200_upload_grains) */
  int32_t lmethod_aniso1;        /* 1: tabulated s11 (Mie), 0: HG */
  int32_t lisotropic;
  int32_t l_sym_centrale, l_sym_axiale;
  int32_t lonly_LTE;             /* init_mcfost.f90:1880-1883; 0 needs mcfost_b200_upload_grains */
  int32_t lxJ_abs_step1;         /* xJ_abs tally during the thermal step */
  int32_t lxJ_abs;               /* xJ_abs tally during SED step */
  /* detectors (read_param.f90:180-184) */
  int32_t N_thet, N_phi, capt_sup;
  /* rt1 observer directions (dust_ray_tracing.f90: tab_u_rt etc.) */
  int32_t RT_n_incl, RT_n_az;
  const double *tab_u_rt;   /* (RT_n_incl, RT_n_az) */
  const double *tab_v_rt;   /* (RT_n_incl, RT_n_az) */
  const double *tab_w_rt;   /* (RT_n_incl)          */
  /* RNG: Philox4x32-10 key; counter = (draw block, packet id, call_index).
   * seed defaults to the reference's 269753 (random_numbers.f90:19). */
  uint64_t seed;
  uint32_t call_index;      /* distinguishes successive calls (iteration / lambda) */
  /* multi-GPU: this process handles chunks c with (c-1) % n_ranks == rank and
   * the running-temperature feedback scales the local tally by n_ranks, like
   * the reference's x nb_proc (thermal_emission.f90:670). */
  int32_t rank, n_ranks;
  int32_t reset_tallies;    /* 1: zero all device tallies before the call */
  /* image step (run_image_mc, dust_transfer.f90:692-824): lmono0 with the rt2 accumulator */
  int32_t loutput_mc;       /* keep the MC photon maps (STOKEI..., output.f90:396-570) in the image step */
  int32_t n_theta_I, n_phi_I;   /* angular bins of I_spec (15 x 15, dust_ray_tracing.f90:104-105) */
  /* grain heating regimes (parameters.f90; lonly_* derived in init_mcfost.f90:1880-1883) */
  int32_t lonly_nLTE, lRE_nLTE, lnRE;
  int32_t low_mem_th_emission_nLTE;   /* 1: select_absorbing_grain instead of kabs_nLTE_CDF */
  int32_t low_mem_scattering;         /* method 1: 1 = on-the-fly CDF, 0 = ksca_CDF (dust_prop.f90:1292) */
  /* capteur, Monte Carlo photon maps and packet origin (output.f90:303-357,396-570): used when lmono0 and
   * loutput_mc (maps), lonly_capt_interet (all modes) and lorigine */
  int32_t npix_x, npix_y;             /* read_param.f90:176 */
  float   zoom;                       /* parameters.f90:109 */
  double  map_size;                   /* parameters.f90:169, AU */
  double  cos_disk, sin_disk;         /* output.f90:94-95: cos / sin(ang_disque) */
  int32_t l_sym_ima;                  /* left-right symmetry: half a photon in each mirror pixel */
  int32_t lonly_capt_interet, capt_inf;   /* keep only detector bins capt_inf..capt_sup (read_param.f90:182-190) */
  int32_t lorigine, capt_interet;     /* star_origin / disk_origin tallies for bin capt_interet */
  /* emission extras (dust_transfer.f90:1090-1142) and the low-memory LTE emission branch */
  int32_t low_mem_th_emission;        /* 1: per-grain kdB_dT_1grain_LTE_CDF + select_absorbing_grain (needs upload_grains) */
  int32_t lweight_emission;           /* disk packets carry correct_E_emission(icell) */
  int32_t lspot;                      /* hot spot on star 1 */
  float   T_spot, surf_fraction_spot, theta_spot, phi_spot;   /* parameters.f90:247 (degrees) */
  double  star1_T;                    /* star(1)%T */
  const double *tab_lambda;           /* (n_lambda) micron, wavelengths.f90:15; read with lspot only */
  /* packet counts per cell (radiation_field.f90:53,60): thermal step with lmcfost_lib (one column), SED / image
   * step with lProDiMo (inside lxJ_abs; one column per wavelength) */
  int32_t lxN_abs;
  /* Modified random walk (MRW.f90, call site dust_transfer.f90:1222-1239 -- disabled and unfinished in the
   * reference; built here to Min et al. 2009 / Robitaille 2010).  Thermal step with lonly_LTE
   * only.  A packet whose last n_iteractions_in_cell > 5 flights all ended in the cell they started in
   * (dust_transfer.f90:1242-1249) takes diffusion steps of radius d = distance_to_closest_wall while
   * d > gamma_MRW * (local Rosseland-type mean free path). */
  int32_t lMRW;
  float   gamma_MRW;                  /* MRW.f90:11 (2.0); <= 0 selects 2.0 */
  /* lProDiMo / lML form of the SED loop (dust_transfer.f90:512-516): chunks end on packets SENT even though the
   * call is not the thermal step, and wavelengths below 0.5 micron send 10x the packets.  The caller passes the
   * resulting per-chunk count in n_photons2 exactly as the Fortran computes it; this flag only selects the
   * termination rule. */
  int32_t lcount_sent;
  /* Packets in flight are capped at max(one warp per SM, max_inflight_fraction * packets sent so far) so that
   * immediate re-emission sees running tallies the way a host run with a few threads does.  <= 0 selects the
   * default (1/32). */
  float   max_inflight_fraction;
  /* The interstellar-radiation-field side loop of run_sed_mc (dust_transfer.f90:941-985, lProDiMo / lML): every packet is
   * emitted by emit_packet_ISM, a chunk ends when n_photons2 packets that ENTER the model were sent (packets that miss it
   * are counted in n_phot_envoyes -- the reference's n_phot_envoyes_ISM -- but not towards the chunk), the ray-tracing
   * accumulators are off.  Not the thermal step. */
  int32_t lISM_loop;
} mcb_run_params;

/* ------------------------------------------------------------------------
 * Tallies returned to the caller (any pointer may be NULL = not wanted).
 * Shapes are the reference's with the trailing nb_proc dimension removed:
 * the shim stores them in the id=1 slice and zeroes the others.
 * --------------------------------------------------------------------- */
typedef struct mcb_tallies {
  double  *xKJ_abs;          /* (n_cells)              radiation_field.f90:20 */
  double  *xJ_abs;           /* (n_cells, n_lambda)    :22 */
  int32_t *xT_ech;           /* (n_cells)              thermal_emission.f90:49 */
  double  *n_phot_envoyes;   /* (n_lambda)             */
  double  *sed, *sed_q, *sed_u, *sed_v, *n_phot_sed;                 /* (n_lambda,N_thet,N_phi) */
  double  *sed_star, *sed_star_scat, *sed_disk, *sed_disk_scat;      /* output.f90:573-589 */
  float   *xI_scatt;         /* (45, 2, N_type_flux, RT_n_incl*RT_n_az, n_cells) real, dust_ray_tracing.f90:33 */
  int32_t  N_type_flux;      /* 1, 4 (pola), 5/8 (contrib) as in the reference */
  /* rt2 accumulators (radiation_field.f90:91-130), `real` like the reference: */
  float   *I_spec;           /* (N_type_flux, n_theta_I, n_phi_I, n_cells)   dust_ray_tracing.f90:44 */
  float   *I_spec_star;      /* (n_cells)                                    dust_ray_tracing.f90:45 */
  /* diagnostics (not in the reference): */
  double  *stats;            /* [12]: packets, cell-steps, interactions, scatterings,
                                absorptions, killed, escaped, dark-zone bounces,
                                MRW walks, MRW steps, 2 spare */
  /* nLTE / qRE (thermal_emission.f90:49,60): in/out like xT_ech when reset_tallies = 0 */
  int32_t *xT_ech_1grain;      /* (grain_RE_nLTE_start:grain_RE_nLTE_end, n_cells) */
  int32_t *xT_ech_1grain_nRE;  /* (grain_nRE_start:grain_nRE_end, n_cells) */
  double  *E_abs_nRE;          /* scalar, dust_transfer.f90:1357 */
  /* Monte Carlo photon maps of ONE wavelength (the call's lambda_in): (npix_x, npix_y, N_thet, N_phi, n_maps)
   * with n_maps = N_type_flux and the planes ordered I [, Q, U, V] [, I_star, I_star_scat, I_disk, I_disk_scat]
   * = STOKEI, STOKEQ, STOKEU, STOKEV, STOKEI_star, ..._disk_scat (lambda,:,:,:,:) of output.f90:26-34 */
  double  *stokes_map;
  double  *star_origin;        /* (n_lambda)           output.f90:37 */
  double  *disk_origin;        /* (n_lambda, n_cells)  output.f90:36 */
  double  *xN_abs;             /* (n_cells) in the thermal step, (n_cells, n_lambda) otherwise; radiation_field.f90:23 */
} mcb_tallies;

/* ---- life cycle -------------------------------------------------------- */
int  mcfost_b200_init(int device, mcb_handle **h);
void mcfost_b200_finalize(mcb_handle *h);
const char *mcfost_b200_last_error(const mcb_handle *h);

int mcfost_b200_upload_grid(mcb_handle *h, const mcb_grid *g);
/* l_dark_zone(n_cells) (cylindrical_grid.f90:38); may change per wavelength
 * in SED mode (dust_transfer.f90:919). NULL = no dark zone. */
int mcfost_b200_upload_dark_zone(mcb_handle *h, const int32_t *l_dark_zone);
int mcfost_b200_upload_opacity(mcb_handle *h, const mcb_opacity *o);
int mcfost_b200_upload_emission(mcb_handle *h, const mcb_emission *e);
/* needs upload_grid and upload_opacity first (extents n_cells, n_lambda, n_T, p_n_cells) */
int mcfost_b200_upload_grains(mcb_handle *h, const mcb_grains *g);

/* The drop-in for mc_photon_loop: blocking; copies tallies D2H into `out`. */
int mcfost_b200_run(mcb_handle *h, const mcb_run_params *r, mcb_tallies *out);

/* Split form used for device-timed benchmarking and multi-GPU reductions:
 * launch only (asynchronous on the handle's stream), expose the packed device
 * tally buffer (fp64 block then fp32 block) so the host layer can all-reduce
 * it in place (one NCCL call), then download. */
int mcfost_b200_launch(mcb_handle *h, const mcb_run_params *r);
int mcfost_b200_sync(mcb_handle *h);
/* Overlap of successive calls (no reference counterpart).  When the packet counter of a thermal call runs
 * dry, ~15 % of the packets in flight belong to the longest-lived 0.1 % (length-biased sampling): ~2 % of the
 * call's events, but up to 3e5 sequential events per packet, i.e. ~1 s during which most SMs idle.  With
 * n_sms_reserved > 0 the main launch of every following call uses all but n_sms_reserved SMs and, once its
 * counter is dry, hands its last packets (<= 128 per SM) to a second launch of n_sms_straggler blocks, so that
 * calls issued on OTHER handles (own streams) can start on the rest of the GPU while this one finishes.
 * With k handles used in turn, n_sms_straggler = n_sms_reserved / (k - 1) keeps every launch resident.
 * Results are unchanged; (0, 0) (default) turns it off.  n_sms_straggler <= n_sms_reserved <= half the SMs;
 * only calls that count packets SENT (thermal, image) hand over. */
int mcfost_b200_set_overlap(mcb_handle *h, int n_sms_reserved, int n_sms_straggler);
int mcfost_b200_tally_buffers(mcb_handle *h, void **d_f64, int64_t *n_f64,
                              void **d_f32, int64_t *n_f32);
int mcfost_b200_download(mcb_handle *h, const mcb_run_params *r, mcb_tallies *out);
/* device time of the last launch in ms (CUDA events on the handle's stream) */
int mcfost_b200_last_kernel_ms(mcb_handle *h, float *ms);
/* scheduling diagnostics of the last launch (not in the reference): out[14] =
 * {ms until the packet counter ran dry, kernel ms, chunk visits[4], valid lanes[4]
 * for the phases EMIT, ABSORB, SCATTER, FLY, packets handed to the straggler launch,
 * ms until the main launch ended, ms until the straggler launch ended,
 * device timer at the start of the main launch in ms, ms until the straggler launch started, kernels launched by the
 * last call}; out must hold 16 doubles */
int mcfost_b200_debug_counters(mcb_handle *h, double *out);
/* cudaStream_t of the handle, as an integer, so torch can wait on it */
int mcfost_b200_stream(mcb_handle *h, uint64_t *stream);

/* ---- post-MC temperature solves on the device-resident tallies of the last call:
 * Temp_finale (thermal_emission.f90:870-906, Temp_LTE with id = 0) -> Tdust(n_cells), `real`;
 * Temp_finale_nLTE (:932-1014) -> Tdust_1grain(grain_RE_nLTE_start:grain_RE_nLTE_end, n_cells), `real`.
 * They read xKJ_abs / xJ_abs / xT_ech* where the photon loop (and the multi-GPU all-reduce) left them, so only
 * the temperatures cross the bus.  Host output pointers. */
int mcfost_b200_temp_finale(mcb_handle *h, float *Tdust);
int mcfost_b200_temp_finale_nlte(mcb_handle *h, float *Tdust_1grain);

/* ---- deterministic sub-kernels (parity tests; also the building blocks of
 * define_dark_zone / integ_tau).  One ray per thread.
 * All arrays are HOST pointers of length n. ------------------------------ */
/* cross_cell (grid.f90:16-22 procedure pointer): one cell crossing */
int mcfost_b200_cross_cell(mcb_handle *h, int64_t n,
        const double *x0, const double *y0, const double *z0,
        const double *u, const double *v, const double *w,
        const int32_t *icell, const int32_t *previous_cell,
        double *x1, double *y1, double *z1, int32_t *next_cell,
        double *l, double *l_contrib, double *l_void_before);
/* index_cell: point location */
int mcfost_b200_index_cell(mcb_handle *h, int64_t n,
        const double *x, const double *y, const double *z, int32_t *icell);
/* move_to_grid: entry from outside; x,y,z updated in place */
int mcfost_b200_move_to_grid(mcb_handle *h, int64_t n,
        double *x, double *y, double *z,
        const double *u, const double *v, const double *w,
        int32_t *icell, int32_t *lintersect);
/* optical_length_tot (optical_depth.f90:248-324): tau to the grid edge along
 * fixed rays at wavelength index lambda (1-based) */
int mcfost_b200_optical_length_tot(mcb_handle *h, int64_t n, int32_t lambda,
        const double *x, const double *y, const double *z,
        const double *u, const double *v, const double *w,
        const int32_t *icell, double *tau_tot, double *lmin, double *lmax,
        int32_t *n_steps);
/* Ray-tracing method 1, the core of the formal solution.
 *
 * init_dust_source_fct1(lambda, ibin, iaz) (dust_ray_tracing.f90:636-708): the
 * source function eps_dust1(n_az_rt, n_theta_rt, N_type_flux, n_cells) from the
 * scattered specific intensity xI_scatt that the last mcfost_b200_run with
 * lscatt_ray_tracing1 LEFT ON THE DEVICE (no download / upload of the tally),
 * for the observer direction iRT = RT2d_to_RT1d(ibin, iaz) (1-based).
 * photon_energy: the caller's (:657-664).  J_th(n_cells): the caller's calc_Jth
 * (thermal emissivity at lambda).  N_type_flux, lsepar_pola, lsepar_contrib are
 * those of that run.  The table stays on the device for
 * mcfost_b200_integ_ray_dust; eps_dust1 (may be NULL) receives a copy with
 * extents (45, 2, N_type_flux, n_cells) (on a 3D grid only (1, 1, :, :) is
 * used, n_az_rt = n_theta_rt = 1).
 *
 * integ_ray_dust(lambda, icell, x, y, z, u, v, w) (optical_depth.f90:1327-1421)
 * with dust_source_fct of method 1 (dust_ray_tracing.f90:1458-1485) for n rays
 * followed backwards from the observer's side: I(N_type_flux, n) = sum over the
 * cells of exp(-tau) (1 - exp(-dtau)) eps_dust1(k(phi), psup(z), :, icell), down
 * to tau_dark_zone_obs.  The pixel loops above it (dust_map, intensite_pixel_dust)
 * stay with the caller. */
int mcfost_b200_init_dust_source_fct1(mcb_handle *h, int32_t lambda, int32_t iRT,
        double photon_energy, const double *J_th, double *eps_dust1);
int mcfost_b200_integ_ray_dust(mcb_handle *h, int32_t lambda, int64_t n,
        const double *x, const double *y, const double *z,
        const double *u, const double *v, const double *w,
        const int32_t *icell, float tau_dark_zone_obs, double *I);
/* init_reemission (thermal_emission.f90:404-550), LTE cells, high-memory branch,
 * no extra heating: the Planck function and its temperature derivative per
 * (lambda, T) with the reference's constants, then per (T, p_icell) the cooling
 * table log(Qcool(T) - Qcool(T_min)) and the emission CDF
 * sum_lambda kappa_abs_LTE dB/dT, computed ON THE DEVICE from the
 * kappa_abs_LTE and tab_Temp of the last mcfost_b200_upload_opacity (whose
 * log_Qcool_minus_extra_heating / kdB_dT_CDF may then be NULL) and installed as
 * the handle's thermal tables: with cell-dependent dust the CDF is n_lambda x
 * n_T x n_cells doubles (280 MB for 7000 cells) that never cross PCIe.
 * tab_lambda, tab_delta_lambda: (n_lambda), micron.  Outputs (may be NULL):
 * log_Qcool_minus_extra_heating(n_T, p_n_cells), kdB_dT_CDF(n_lambda, n_T,
 * p_n_cells) for the caller's own use (Temp_finale on the host, FITS output). */
int mcfost_b200_init_reemission(mcb_handle *h, const double *tab_lambda,
        const double *tab_delta_lambda,
        double *log_Qcool_minus_extra_heating, double *kdB_dT_CDF);
/* The per-grain tables of init_reemission for grains k_start..k_end (1-based,
 * inclusive) of C_abs_norm(n_grains_tot, n_lambda) (`real`):
 * log_E_em_1grain(k_start:k_end, n_T) (:551-567 for the nLTE grains, :585-603
 * for the nRE grains, whose E_em_1grain_nRE is the optional E_em_1grain) and
 * kdB_dT_1grain_*_CDF(n_lambda, k_start:k_end, n_T) (:569-581, :605-618, also
 * the low-memory LTE table :519-533).  Pure function of its arguments and of
 * the handle's n_lambda / n_T / tab_Temp; the caller passes the results on in
 * mcb_grains. */
int mcfost_b200_init_reemission_grains(mcb_handle *h, const double *tab_lambda,
        const double *tab_delta_lambda, const float *C_abs_norm,
        int32_t n_grains_tot, int32_t k_start, int32_t k_end,
        double *log_E_em_1grain, double *E_em_1grain, double *kdB_dT_1grain_CDF);
/* repartition_energie(lambda) (thermal_emission.f90:1771-1949), LTE case, for
 * lambda_first..lambda_last (1-based, inclusive), on the device: the thermal
 * emission of every cell from Tdust(n_cells) (`real`) with the reference's
 * constants, E_disk(lambda) = its sum, prob_E_cell(0:n_cells, lambda) = its
 * cumulative sum (times weight_proba_emission(n_cells) when given,
 * lweight_emission), normalised; dark cells (the handle's dark zone) do not
 * emit.  The sums are parallel reductions / scans (1e-12 from the sequential
 * sums of the reference).  tab_lambda(n_lambda) in micron; E_stars, E_ISM
 * (n_lambda; E_ISM may be NULL).  Outputs for the wavelengths of the range:
 * E_disk, frac_E_stars, frac_E_disk (n_lambda arrays), weight_norm (may be
 * NULL: prob_E_cell(n_cells) / E_disk before normalisation, the factor of
 * :1932-1934), prob_E_cell (may be NULL: a host copy of the columns).
 * prob_E_cell, frac_E_stars and frac_E_disk STAY ON THE DEVICE as the handle's
 * emission tables: a following mcfost_b200_upload_emission may pass NULL for
 * those three (with 1e6 cells prob_E_cell is 8 MB per wavelength that is
 * neither built nor uploaded by the host).  E_totale and repartition_wl_em
 * (n_lambda-long loops on E_disk) stay with the caller.  Returns an error when
 * a wavelength has no energy at all (the reference exits, :1900-1904). */
int mcfost_b200_repartition_energie(mcb_handle *h, int32_t lambda_first, int32_t lambda_last,
        const float *Tdust, const double *tab_lambda,
        const double *E_stars, const double *E_ISM, const double *weight_proba_emission,
        double *E_disk, double *frac_E_stars, double *frac_E_disk, double *weight_norm,
        double *prob_E_cell);
/* define_dark_zone(lambda, p_lambda, tau_max, ldiff_approx) (optical_depth.f90:
 * 1425-1651) on structured grids: radial and vertical optical-depth sums
 * (`real` accumulators as in the Fortran), then 11 rays of optical depth
 * tau_max from the centre of every candidate cell, column after column, each
 * column seeing the cells the previous ones made dark.  r_grid, z_grid:
 * (n_cells) cell centres.  regions: (iRmin, iRmax) of regions(:), whose edge
 * columns are never dark (:1640-1645).  dust_density_sum: NULL, or
 * sum(dust_density_o_n_grains(:,icell)) when n_zones > 1 (:1634-1638).
 * Outputs: l_dark_zone(n_cells) as int32, ri_in / ri_out_dark_zone(n_az),
 * l_is_dark_zone; zj_sup / zj_inf_dark_zone(n_rad, n_az) are IN-OUT (module
 * arrays that keep their values between calls, zero-initialised in mem.f90:
 * 162-170; zj_inf may be NULL on a 2D grid).  The `first cell is in diffusion
 * approximation zone` check (:1629-1632) is the caller's, on ri_in_dark_zone.
 * The result is also installed as the handle's dark zone (as
 * mcfost_b200_upload_dark_zone would).  Kept as in the reference: the lower-
 * half loop of the 3D branch marks nothing (:1607-1610 runs zero times). */
int mcfost_b200_define_dark_zone(mcb_handle *h, int32_t lambda, float tau_max,
        const double *r_grid, const double *z_grid,
        int32_t n_regions, const int32_t *region_iRmin, const int32_t *region_iRmax,
        const double *dust_density_sum,
        int32_t *l_dark_zone, int32_t *ri_in_dark_zone, int32_t *ri_out_dark_zone,
        int32_t *zj_sup_dark_zone, int32_t *zj_inf_dark_zone, int32_t *l_is_dark_zone);
/* compute_column (optical_depth.f90:328-415): from the centre of every cell,
 * along 4 directions (1: towards the star at the origin, 2: +z, 3: -z, 4:
 * radially outwards), the sum of l_contrib * factor over the cells crossed.
 * factor == NULL: the optical depth at wavelength index lambda (type 2:
 * kappa(lambda) * kappa_factor); else factor[n_cells] is the caller's per-cell
 * weight (types 1 and 3: CD_units * gas_density [* tab_abundance]) and lambda
 * is ignored.  centre_x/y/z[n_cells]: the caller's r_grid cos(phi_grid),
 * r_grid sin(phi_grid), z_grid (or the Voronoi seeds), :362-370.
 * column: real(n_cells, 4), column-major.  With lvariable_dust the opacity of a
 * cell is the cell's own (the reference reads the previous cell's, :391-393). */
int mcfost_b200_compute_column(mcb_handle *h, int32_t lambda, const double *factor,
        const double *centre_x, const double *centre_y, const double *centre_z,
        float *column);
/* physical_length (optical_depth.f90:21-182) with Stokes = 0 (no tallies), as
 * define_dark_zone uses it (optical_depth.f90:1536): walk to optical depth
 * tau. x,y,z,u,v,w,icell updated in place like the reference's inout args. */
int mcfost_b200_physical_length(mcb_handle *h, int64_t n, int32_t lambda,
        double *x, double *y, double *z, double *u, double *v, double *w,
        int32_t *icell, const float *tau, float *ltot,
        int32_t *flag_sortie, int32_t *lpacket_alive);

/* distance_to_closest_wall (grid.f90 procedure pointer -> cylindrical_grid.f90:1179-1226,
 * spherical_grid.f90:451-499, Voronoi.f90:996-1061): radius of the largest sphere around (x,y,z)
 * that stays inside cell icell; the modified random walk steps by it. */
int mcfost_b200_distance_to_closest_wall(mcb_handle *h, int64_t n, const int32_t *icell,
        const double *x, const double *y, const double *z, double *s);
/* Mean opacities of the modified random walk per temperature index (compute_Planck_opacities,
 * diffusion.f90:631-693, restated for the running re-emission spectrum kdB_dT_CDF):
 * A, B, C are host arrays (n_T, p_n_cells).  Built on the device at the first lMRW call or here. */
int mcfost_b200_mrw_tables(mcb_handle *h, double *A, double *B, double *C);

/* ---- multi-GPU behind the boundary: ONE host process, the n GPUs of a node.  Grid and tables are
 * replicated (the multi_upload_* calls upload to every GPU), GPU g runs the chunks c with (c-1) mod n == g, and
 * mcfost_b200_multi_run merges EVERY tally of the call with one group of NCCL all-reduces over NVLink before it
 * returns the merged tallies of GPU 0: sum for the packed fp64 block, xI_scatt, I_spec, I_spec_star, the photon maps,
 * the origin tallies and xN_abs (sum(...,dim=id) in the reference: thermal_emission.f90:668, output.f90:3084-3102,
 * dust_ray_tracing.f90:661,689); min for xT_ech (Temp_LTE with id = 0 starts from minval(xT_ech(icell,:)),
 * thermal_emission.f90:683); max for xT_ech_1grain / xT_ech_1grain_nRE (maxval, :823, :977).  r->rank / r->n_ranks are
 * set by the library.  NCCL (libnccl.so.2) is loaded at run time when n_gpus > 1. */
typedef struct mcb_multi mcb_multi;
int  mcfost_b200_multi_init(int n_gpus, const int *devices /* NULL: 0..n_gpus-1 */, mcb_multi **m);
void mcfost_b200_multi_finalize(mcb_multi *m);
const char *mcfost_b200_multi_last_error(const mcb_multi *m);
int  mcfost_b200_multi_n_gpus(const mcb_multi *m);
mcb_handle *mcfost_b200_multi_handle(mcb_multi *m, int i);      /* the per-GPU handle (deterministic kernels, diagnostics) */
int mcfost_b200_multi_upload_grid(mcb_multi *m, const mcb_grid *g);
int mcfost_b200_multi_upload_dark_zone(mcb_multi *m, const int32_t *l_dark_zone);
int mcfost_b200_multi_upload_opacity(mcb_multi *m, const mcb_opacity *o);
int mcfost_b200_multi_upload_emission(mcb_multi *m, const mcb_emission *e);
int mcfost_b200_multi_upload_grains(mcb_multi *m, const mcb_grains *g);
int mcfost_b200_multi_run(mcb_multi *m, const mcb_run_params *r, mcb_tallies *out);
int mcfost_b200_multi_temp_finale(mcb_multi *m, float *Tdust);

#ifdef __cplusplus
}
#endif
#endif /* MCFOST_B200_H */
