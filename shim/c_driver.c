/*
 * A plain-C consumer of include/mcfost_b200.h: fills every struct of the ABI field by field, in the header's order,
 * the way the Fortran shim (shim/mcfost_b200_shim.f90) does with c_loc() on MCFOST's module arrays, and runs one thermal
 * mc_photon_loop call (dust_transfer.f90:439-572) through libmcfost_b200.so.  The arrays come from a flat binary dump of
 * a synthetic problem (tools/dump_problem.py) instead of MCFOST's setup routines.
 *
 *   gcc -std=c99 -O2 -I include -o c_driver shim/c_driver.c -L mcfost_b200/_lib -lmcfost_b200 -Wl,-rpath,$PWD/mcfost_b200/_lib
 *   ./c_driver problem.bin [n_photons2]
 * Exit code 0 = the call ran and conserved its packets; 77 = no CUDA device (MCB_ERR_NO_DEVICE: there is no CPU fallback).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mcfost_b200.h"

typedef struct { char name[33]; int code; long long n; void *data; } rec_t;
static rec_t recs[128];
static int n_recs = 0;

static int load(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  for (;;) {
    char name[32]; int code; long long n;
    if (fread(name, 1, 32, f) != 32) break;
    if (fread(&code, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) { fclose(f); return -2; }
    size_t sz = (code == 0 ? 8 : 4) * (size_t)n;
    rec_t *r = &recs[n_recs++];
    memcpy(r->name, name, 32); r->name[32] = 0; r->code = code; r->n = n; r->data = malloc(sz ? sz : 1);
    if (fread(r->data, 1, sz, f) != sz) { fclose(f); return -3; }
    if (n_recs == 128) break;
  }
  fclose(f);
  return 0;
}
static void *arr(const char *name) { for (int i = 0; i < n_recs; ++i) if (!strcmp(recs[i].name, name)) return recs[i].data; return NULL; }
static int geti(const char *name) { int *p = (int *)arr(name); return p ? *p : 0; }
static double getd(const char *name) { double *p = (double *)arr(name); return p ? *p : 0.0; }

#define CHECK(call, what) do { int rc_ = (call); if (rc_) { fprintf(stderr, "mcfost_b200: %s failed (%d): %s\n", what, rc_, mcfost_b200_last_error(h)); return rc_ == MCB_ERR_NO_DEVICE ? 77 : 1; } } while (0)

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s problem.bin [n_photons2]\n", argv[0]); return 2; }
  if (load(argv[1])) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const int n_photons2 = argc > 2 ? atoi(argv[2]) : 100;
  mcb_handle *h = NULL;
  CHECK(mcfost_b200_init(0, &h), "init");

  /* ---- mcb_grid: cylindrical_grid.f90:20-41 + star(:) ---- */
  mcb_grid g; memset(&g, 0, sizeof g);
  g.kind = geti("kind"); g.l3D = geti("l3D");
  g.n_rad = geti("n_rad"); g.nz = geti("nz"); g.n_az = geti("n_az");
  g.n_cells = geti("n_cells");
  g.Rmax2 = getd("Rmax2"); g.zmaxmax = getd("zmaxmax");
  g.r_lim = arr("r_lim"); g.r_lim_2 = arr("r_lim_2"); g.r_lim_3 = arr("r_lim_3");
  g.z_lim = arr("z_lim"); g.zmax = arr("zmax");
  g.tan_theta_lim = NULL; g.theta_lim = NULL;          /* spherical grids only */
  g.tan_phi_lim = arr("tan_phi_lim"); g.volume = arr("volume");
  g.n_cells_tot = geti("n_cells_tot");
  g.cell_map_i = arr("cell_map_i"); g.cell_map_j = arr("cell_map_j"); g.cell_map_k = arr("cell_map_k");
  /* Voronoi members stay NULL / 0 on a cylindrical grid */
  g.n_stars = geti("n_stars");
  g.star_xyzr = arr("star_xyzr"); g.star_icell = arr("star_icell"); g.star_out_model = arr("star_out_model");
  g.w_lim = NULL; g.sin_phi_lim = NULL; g.cos_phi_lim = NULL;      /* distance_to_closest_wall tables: spherical / 3D only */
  CHECK(mcfost_b200_upload_grid(h, &g), "upload_grid");

  /* ---- mcb_opacity: dust_prop.f90:17-21, thermal_emission.f90:404-644 ---- */
  mcb_opacity o; memset(&o, 0, sizeof o);
  o.n_lambda = geti("n_lambda"); o.p_n_cells = geti("p_n_cells"); o.p_n_lambda_pos = geti("p_n_lambda_pos"); o.n_T = geti("n_T");
  o.kappa = arr("kappa"); o.kappa_abs_LTE = arr("kappa_abs_LTE"); o.kappa_factor = arr("kappa_factor");
  o.tab_albedo_pos = arr("tab_albedo_pos"); o.tab_g_pos = arr("tab_g_pos");
  o.prob_s11_pos = arr("prob_s11_pos"); o.tab_s11_pos = arr("tab_s11_pos");
  o.tab_s12_o_s11_pos = arr("tab_s12_o_s11_pos"); o.tab_s22_o_s11_pos = arr("tab_s22_o_s11_pos"); o.tab_s33_o_s11_pos = arr("tab_s33_o_s11_pos");
  o.tab_s34_o_s11_pos = arr("tab_s34_o_s11_pos"); o.tab_s44_o_s11_pos = arr("tab_s44_o_s11_pos");
  o.log_Qcool_minus_extra_heating = arr("log_Qcool_minus_extra_heating"); o.kdB_dT_CDF = arr("kdB_dT_CDF");
  o.tab_Temp = arr("tab_Temp"); o.T_min = (float)getd("T_min");
  CHECK(mcfost_b200_upload_opacity(h, &o), "upload_opacity");
  CHECK(mcfost_b200_upload_dark_zone(h, (const int32_t *)arr("l_dark_zone")), "upload_dark_zone");

  /* ---- mcb_emission: repartition_energie, thermal_emission.f90:1771 ---- */
  mcb_emission e; memset(&e, 0, sizeof e);
  e.spectre_emission_cumul = arr("spectre_emission_cumul"); e.frac_E_stars = arr("frac_E_stars"); e.frac_E_disk = arr("frac_E_disk");
  e.prob_E_cell = arr("prob_E_cell"); e.CDF_E_star = arr("CDF_E_star");
  e.L_packet_th = getd("L_packet_th") * 100.0 / n_photons2;      /* the dump's L_packet_th is for 100 packets per chunk */
  e.E_paquet = getd("E_paquet"); e.R_ISM = 0.0; e.centre_ISM[0] = e.centre_ISM[1] = e.centre_ISM[2] = 0.0;
  e.correct_E_emission = NULL;
  CHECK(mcfost_b200_upload_emission(h, &e), "upload_emission");

  /* ---- mcb_run_params: the six dummy arguments of mc_photon_loop, then the module-level flags ---- */
  mcb_run_params r; memset(&r, 0, sizeof r);
  r.lambda_in = 1; r.p_lambda_in = 1; r.n_photons2 = n_photons2; r.n_phot_lim = 1.0e30f; r.nnfot1_start = 1; r.laffichage = 0;
  r.n_photons_loop = 128;
  r.letape_th = 1; r.lmono = 0; r.lmono0 = 0;
  r.lscatt_ray_tracing1 = 0; r.lscatt_ray_tracing2 = 0;
  r.lsepar_pola = 1; r.lsepar_contrib = 1;
  r.lscattering_method1 = 0; r.lmethod_aniso1 = 1; r.lisotropic = 0;
  r.l_sym_centrale = 1; r.l_sym_axiale = 1;
  r.lonly_LTE = 1; r.lxJ_abs_step1 = 0; r.lxJ_abs = 0;
  r.N_thet = 10; r.N_phi = 1; r.capt_sup = 2;
  r.RT_n_incl = 0; r.RT_n_az = 0; r.tab_u_rt = NULL; r.tab_v_rt = NULL; r.tab_w_rt = NULL;
  r.seed = 269753; r.call_index = 0;
  r.rank = 0; r.n_ranks = 1; r.reset_tallies = 1;
  r.loutput_mc = 0; r.n_theta_I = 15; r.n_phi_I = 15;
  r.lonly_nLTE = 0; r.lRE_nLTE = 0; r.lnRE = 0; r.low_mem_th_emission_nLTE = 0; r.low_mem_scattering = 1;
  r.npix_x = 0; r.npix_y = 0; r.zoom = 1.0f; r.map_size = 0.0; r.cos_disk = 1.0; r.sin_disk = 0.0; r.l_sym_ima = 0;
  r.lonly_capt_interet = 0; r.capt_inf = 1; r.lorigine = 0; r.capt_interet = 1;
  r.low_mem_th_emission = 0; r.lweight_emission = 0; r.lspot = 0;
  r.T_spot = 0.0f; r.surf_fraction_spot = 0.0f; r.theta_spot = 0.0f; r.phi_spot = 0.0f; r.star1_T = 0.0; r.tab_lambda = NULL;
  r.lxN_abs = 0;
  r.lMRW = 0; r.gamma_MRW = 2.0f; r.lcount_sent = 0; r.max_inflight_fraction = 0.0f;
  r.lISM_loop = 0;

  /* ---- mcb_tallies: caller-allocated, the id = 1 slices of the reference's (..., nb_proc) arrays ---- */
  const size_t nc = (size_t)g.n_cells, nl = (size_t)o.n_lambda, nsed = nl * (size_t)r.N_thet * (size_t)r.N_phi;
  mcb_tallies t; memset(&t, 0, sizeof t);
  t.xKJ_abs = calloc(nc, 8); t.xJ_abs = NULL; t.xT_ech = calloc(nc, 4); t.n_phot_envoyes = calloc(nl, 8);
  t.sed = calloc(nsed, 8); t.sed_q = calloc(nsed, 8); t.sed_u = calloc(nsed, 8); t.sed_v = calloc(nsed, 8); t.n_phot_sed = calloc(nsed, 8);
  t.sed_star = calloc(nsed, 8); t.sed_star_scat = calloc(nsed, 8); t.sed_disk = calloc(nsed, 8); t.sed_disk_scat = calloc(nsed, 8);
  t.stats = calloc(12, 8);
  CHECK(mcfost_b200_run(h, &r, &t), "run");

  double sent = 0, sed = 0, xkj = 0;
  for (size_t l = 0; l < nl; ++l) sent += t.n_phot_envoyes[l];
  for (size_t i = 0; i < nsed; ++i) sed += t.sed[i];
  for (size_t i = 0; i < nc; ++i) xkj += t.xKJ_abs[i];
  printf("packets %.0f  sent %.0f  escaped %.0f  killed %.0f  sum(sed) %.6f  sum(xKJ_abs) %.6e  steps %.0f\n", t.stats[0], sent, t.stats[6], t.stats[5], sed, xkj, t.stats[1]);
  int ok = t.stats[0] == 128.0 * n_photons2 && sent == t.stats[0] && t.stats[5] + t.stats[6] == t.stats[0] && xkj > 0;
  float *Tdust = calloc(nc, 4);
  CHECK(mcfost_b200_temp_finale(h, Tdust), "temp_finale");
  float tmax = 0; for (size_t i = 0; i < nc; ++i) if (Tdust[i] > tmax) tmax = Tdust[i];
  printf("Tdust max %.2f K\n", tmax);

  /* ---- the neighbours of the path a caller binds the same way ---- */
  /* define_dark_zone(lambda, p_lambda, tau_max, ldiff_approx), optical_depth.f90:1425-1651 */
  const int lam = geti("lambda_seuil");
  int32_t *dark = calloc(nc, 4), *zj_sup = calloc((size_t)g.n_rad, 4), ri_in = 0, ri_out = 0, is_dark = 0;
  int32_t iRmin = 1, iRmax = g.n_rad;
  CHECK(mcfost_b200_define_dark_zone(h, lam, 30.0f, arr("r_grid"), arr("z_grid"), 1, &iRmin, &iRmax, NULL, dark, &ri_in, &ri_out, zj_sup, NULL, &is_dark),
        "define_dark_zone");
  long n_dark = 0; for (size_t i = 0; i < nc; ++i) n_dark += dark[i];
  printf("dark zone: %ld cells, ri_in %d ri_out %d l_is_dark_zone %d\n", n_dark, ri_in, ri_out, is_dark);
  /* compute_column(2, column, lambda), optical_depth.f90:328-415 (2D grid: the cell centres are (r_grid, 0, z_grid)) */
  double *cy = calloc(nc, 8);
  float *column = calloc(4 * nc, 4);
  CHECK(mcfost_b200_compute_column(h, lam, NULL, arr("r_grid"), cy, arr("z_grid"), column), "compute_column");
  float cmax = 0; for (size_t i = 0; i < 4 * nc; ++i) if (column[i] > cmax) cmax = column[i];
  printf("column: max optical depth %.4e\n", cmax);
  /* init_reemission, thermal_emission.f90:404-550: the tables on the device against the uploaded ones */
  double *logQ = calloc((size_t)o.n_T * o.p_n_cells, 8);
  CHECK(mcfost_b200_init_reemission(h, arr("tab_lambda"), arr("tab_delta_lambda"), logQ, NULL), "init_reemission");
  double dq = 0; for (int k = 1; k < o.n_T * o.p_n_cells; ++k) { double d = logQ[k] - o.log_Qcool_minus_extra_heating[k]; if (d < 0) d = -d; if (d > dq) dq = d; }
  printf("init_reemission: max |log Qcool - uploaded| %.3e\n", dq);
  mcfost_b200_finalize(h);
  return ok && tmax > 10.0f && n_dark > 0 && is_dark == 1 && cmax > 100.0f && dq < 1.0e-6 ? 0 : 1;
}
