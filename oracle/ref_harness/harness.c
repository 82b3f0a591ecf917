/* oracle/ref_harness/harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin C glue that exposes the UNMODIFIED reference CPU hot path (compiled from
 * its own sources under /root/reference by oracle/build_ref.sh into
 * oracle/_ref/libmilcref*.so) through a ctypes-friendly ABI.  It plays the role
 * an application's control.c/setup.c plays in MILC:
 *   - layout + neighbour tables  (ks_spectrum/setup.c:30-61, 1405-1445)
 *   - fn_links_t filled from caller-provided fat/long arrays
 *     (generic_ks/fn_links_milc.c:245-302) + back links for the dblstore dslash
 *   - dslash_fn_field            (generic_ks/dslash_fn_dblstore.c:285-304)
 *   - ks_congrad_parity_cpu      (generic_ks/d_congrad5_fn_milc.c:60-407)
 *   - ks_multicg_offset_field_cpu(generic_ks/ks_multicg_offset.c:63-505)
 * Used only by tests/, bench.py's cpu_baseline / --impl reference legs and the
 * golden-vector generator.  Nothing in the product links or loads it.
 */
#define CONTROL
#include "generic_ks_includes.h"
#include "../include/fn_links.h"
#include "../include/ks_action_paths.h"
#include "../include/fermion_links_milc.h"
#include "../include/info.h"
#include "../include/su3_mat_op.h"
#include <string.h>

static fn_links_t *h_fn = NULL;
static int h_ready = 0;

/* Third-neighbour map: same arithmetic as every MILC KS application's setup.c
   (ks_spectrum/setup.c:1427-1445).  It is static there, so the harness has to
   supply its own. */
static void h_third_neighbor(int x, int y, int z, int t, int *dirpt, int FB,
                             int *xp, int *yp, int *zp, int *tp) {
  int dir = (FB == FORWARDS) ? *dirpt : OPP_DIR(*dirpt);
  *xp = x; *yp = y; *zp = z; *tp = t;
  switch (dir) {
  case XUP:   *xp = (x + 3) % nx; break;
  case XDOWN: *xp = (x + 4 * nx - 3) % nx; break;
  case YUP:   *yp = (y + 3) % ny; break;
  case YDOWN: *yp = (y + 4 * ny - 3) % ny; break;
  case ZUP:   *zp = (z + 3) % nz; break;
  case ZDOWN: *zp = (z + 4 * nz - 3) % nz; break;
  case TUP:   *tp = (t + 3) % nt; break;
  case TDOWN: *tp = (t + 4 * nt - 3) % nt; break;
  default: printf("h_third_neighbor: bad direction\n"); exit(1);
  }
}

fn_links_t *milcref_fn(void) { return h_fn; }   /* eigcg_harness.c */
int milcref_precision(void) { return MILC_PRECISION; }
int milcref_sizeof_real(void) { return (int)sizeof(Real); }

/* One lattice geometry per process (MILC's globals are process-wide). */
int milcref_init(int lx, int ly, int lz, int lt) {
  int i;
  if (h_ready) {
    if (lx == nx && ly == ny && lz == nz && lt == nt) return 0;
    return -1;
  }
  initialize_machine(NULL, NULL);
  nx = lx; ny = ly; nz = lz; nt = lt;
  volume = nx * ny * nz * nt;
  this_node = mynode();
  number_of_nodes = numnodes();
  total_iters = 0;
  phases_in = 1;
  setup_layout();
  make_lattice();
  make_nn_gathers();
  for (i = XUP; i <= TUP; i++)
    make_gather(h_third_neighbor, &i, WANT_INVERSE, ALLOW_EVEN_ODD, SWITCH_PARITY);
  sort_eight_gathers(X3UP);
  h_ready = 1;
  return 0;
}

long milcref_sites_on_node(void) { return (long)sites_on_node; }
int milcref_node_index(int x, int y, int z, int t) { return (int)node_index(x, y, z, t); }
int milcref_total_iters(void) { return total_iters; }

/* fat, lng: su3_matrix[4*sites_on_node] in MILC order (fat[4*i+dir]), Real precision. */
int milcref_set_links(const Real *fat, const Real *lng, double eps_naik) {
  if (!h_ready) return -1;
  if (h_fn != NULL) destroy_fn_links(h_fn);
  h_fn = create_fn_links();
  memcpy(h_fn->fat, fat, sizeof(su3_matrix) * 4 * sites_on_node);
  memcpy(h_fn->lng, lng, sizeof(su3_matrix) * 4 * sites_on_node);
  h_fn->eps_naik = eps_naik;
#ifdef DBLSTORE_FN
  load_fn_backlinks(h_fn);
#endif
  return 0;
}

void milcref_dslash(const Real *src, Real *dest, int parity) {
  dslash_fn_field((su3_vector *)src, (su3_vector *)dest, parity, h_fn);
}

/* out[0..6] = final_rsq, final_relrsq, size_r, size_relr, final_iters, final_restart, converged */
static void h_unpack(const quark_invert_control *qic, double *out) {
  out[0] = qic->final_rsq; out[1] = qic->final_relrsq;
  out[2] = qic->size_r;    out[3] = qic->size_relr;
  out[4] = qic->final_iters; out[5] = qic->final_restart; out[6] = qic->converged;
}

int milcref_congrad(const Real *src, Real *dest, double m, int parity,
                    int max, int nrest, double resid, double relresid, double *out) {
  quark_invert_control qic;
  int it;
  memset(&qic, 0, sizeof(qic));
  qic.prec = MILC_PRECISION; qic.min = 0; qic.max = max; qic.nrestart = nrest;
  qic.parity = parity; qic.start_flag = 1; qic.nsrc = 1;
  qic.resid = resid; qic.relresid = relresid;
  it = ks_congrad_parity_cpu((su3_vector *)src, (su3_vector *)dest, &qic, (Real)m, h_fn);
  h_unpack(&qic, out);
  return it;
}

/* psim: num_offsets contiguous fields of sites_on_node su3_vectors; out: 7 doubles per offset */
int milcref_multicg(const Real *src, Real *psim, const double *offsets, int num_offsets,
                    int parity, int max, int nrest, double resid, double relresid,
                    double *out) {
  quark_invert_control *qic = (quark_invert_control *)calloc(num_offsets, sizeof(*qic));
  ks_param *ksp = (ks_param *)calloc(num_offsets, sizeof(*ksp));
  su3_vector **pp = (su3_vector **)malloc(num_offsets * sizeof(*pp));
  int j, it;
  for (j = 0; j < num_offsets; j++) {
    qic[j].prec = MILC_PRECISION; qic[j].max = max; qic[j].nrestart = nrest;
    qic[j].parity = parity; qic[j].nsrc = 1;
    qic[j].resid = resid; qic[j].relresid = relresid;
    ksp[j].offset = offsets[j];
    pp[j] = (su3_vector *)psim + (size_t)j * sites_on_node;
  }
  it = ks_multicg_offset_field_cpu((su3_vector *)src, pp, ksp, num_offsets, qic, h_fn);
  for (j = 0; j < num_offsets; j++) h_unpack(&qic[j], out + 7 * j);
  free(qic); free(ksp); free(pp);
  return it;
}

/* Back-to-back dslash timing for the CPU baseline.  Returns seconds per
   dslash_fn_field call on one parity (src is read on the other parity). */
double milcref_time_dslash(const Real *src, Real *dest, int parity, int ncalls) {
  int k;
  double t0;
  dslash_fn_field((su3_vector *)src, (su3_vector *)dest, parity, h_fn); /* warm */
  t0 = dclock();
  for (k = 0; k < ncalls; k++)
    dslash_fn_field((su3_vector *)src, (su3_vector *)dest, parity, h_fn);
  return (dclock() - t0) / ncalls;
}

/* ---- HISQ link construction --------------------------------------
 * The reference's own chain, create_hisq_links_milc (generic_ks/fermion_links_hisq_load_milc.c:
 * 684-713): U -> V (fat7, load_fatlinks_cpu) -> Y = W (U(3) projection, u3_unitarize_analytic)
 * -> fat, long (asqtad-like smear of W + Naik, load_fatlinks_cpu + load_lnglinks), with the
 * reference's own path tables and coefficients (ks_action_paths_hisq.c, hisq_u3_action.h).
 * links: su3_matrix[4*sites_on_node], KS phases and boundary signs already in (phases_in = 1).
 * Outputs (each su3_matrix[4*sites_on_node], any may be NULL): V, W, fat, lng.
 * coeffs[18]: {one_link, naik, three_staple, five_staple, seven_staple, lepage} of p1, p2, p3.
 * Returns the number of links that took the SVD branch. */
static ks_action_paths_hisq *h_ap = NULL;

static void h_coeffs(double *c, const asqtad_coeffs_t *a) {
  c[0] = a->one_link; c[1] = a->naik; c[2] = a->three_staple;
  c[3] = a->five_staple; c[4] = a->seven_staple; c[5] = a->lepage;
}

int milcref_hisq_links(const Real *links, Real *V, Real *W, Real *fat, Real *lng, double *coeffs) {
  info_t info = INFO_ZERO;
  fn_links_t *fn[1] = {NULL};
  fn_links_t *fn_deps = NULL;
  hisq_auxiliary_t *aux = NULL;
  double eps[1] = {0.0};
  size_t bytes = sizeof(su3_matrix) * 4 * sites_on_node;
  int nsvd;
  if (!h_ready) return -1;
  if (h_ap == NULL) {
    h_ap = create_path_table_hisq();
    make_path_table_hisq(h_ap, 1, eps);
  }
  create_hisq_links_milc(&info, fn, &fn_deps, &aux, h_ap, (su3_matrix *)links, 0, 0);
  nsvd = INFO_HISQ_SVD_COUNTER(&info);
  if (V) memcpy(V, aux->V_link, bytes);
  if (W) memcpy(W, aux->W_unitlink, bytes);
  if (fat) memcpy(fat, get_fatlinks(fn[0]), bytes);
  if (lng) memcpy(lng, get_lnglinks(fn[0]), bytes);
  if (coeffs) {
    h_coeffs(coeffs, &h_ap->p1.act_path_coeff);
    h_coeffs(coeffs + 6, &h_ap->p2.act_path_coeff);
    h_coeffs(coeffs + 12, &h_ap->p3.act_path_coeff);
  }
  destroy_hisq_links_milc(h_ap, aux, fn, fn_deps);
  return nsvd;
}

/* One smearing level on its own: fat (and lng if not NULL) from `links` with the given six
 * coefficients -- load_fatlinks_cpu + load_lnglinks (generic_ks/fermion_links_fn_load_milc.c:
 * 45-107,120-275) with the level-2 path table (the Naik paths carry coeffs[1]). */
int milcref_smear(const Real *links, const double *coeffs, Real *fat, Real *lng) {
  info_t info = INFO_ZERO;
  ks_component_paths p;
  double eps[1] = {0.0};
  int k;
  if (!h_ready) return -1;
  if (h_ap == NULL) {
    h_ap = create_path_table_hisq();
    make_path_table_hisq(h_ap, 1, eps);
  }
  p = h_ap->p2;
  p.act_path_coeff.one_link = coeffs[0]; p.act_path_coeff.naik = coeffs[1];
  p.act_path_coeff.three_staple = coeffs[2]; p.act_path_coeff.five_staple = coeffs[3];
  p.act_path_coeff.seven_staple = coeffs[4]; p.act_path_coeff.lepage = coeffs[5];
  load_fatlinks_cpu(&info, (su3_matrix *)fat, &p, (su3_matrix *)links);
  if (lng) {
    /* load_lnglinks reads the coefficient from the path table: scale a copy of the Naik paths */
    Q_path *q = (Q_path *)malloc(sizeof(Q_path) * p.num_q_paths);
    double naik0 = h_ap->p2.act_path_coeff.naik;
    memcpy(q, p.q_paths, sizeof(Q_path) * p.num_q_paths);
    for (k = 0; k < p.num_q_paths; k++)
      if (q[k].length == 3 && q[k].dir[0] == q[k].dir[1] && q[k].dir[1] == q[k].dir[2])
        q[k].coeff *= (Real)(coeffs[1] / naik0);
    p.q_paths = q;
    load_lnglinks(&info, (su3_matrix *)lng, &p, (su3_matrix *)links);
    free(q);
  }
  return 0;
}

/* W = U(3) projection of every link of V (u3_unitarize_analytic, generic_ks/su3_mat_op.c:828-1205,
 * with the SVD branch as compiled: -DHISQ_REUNIT_ALLOW_SVD, thresholds 1e-8).  Returns SVD count. */
int milcref_unitarize(const Real *V, Real *W, long nlinks) {
  info_t info = INFO_ZERO;
  long k;
  for (k = 0; k < nlinks; k++)
    u3_unitarize_analytic(&info, (su3_matrix *)V + k, (su3_matrix *)W + k);
  return INFO_HISQ_SVD_COUNTER(&info);
}

/* ---- UML propagator solve -----------------------------------------
 * mat_invert_uml_field / mat_invert_block_uml, generic_ks/mat_invert.c:328-402,409-475: dst =
 * (D + 2m)^-1 src on all sites.  src/dst: nsrc contiguous fields of sites_on_node su3_vectors
 * (dst = initial guess in, solution out).  out: 7 doubles (qic after the sequence). */
int milcref_mat_invert_uml(const Real *src, Real *dst, int nsrc, double m, int max, int nrest, double resid,
                           double *out) {
  quark_invert_control qic;
  su3_vector **s = (su3_vector **)malloc(nsrc * sizeof(*s)), **d = (su3_vector **)malloc(nsrc * sizeof(*d));
  int k, it;
  memset(&qic, 0, sizeof(qic));
  param.eigen_param.Nvecs = 0;
  qic.prec = MILC_PRECISION; qic.min = 0; qic.max = max; qic.nrestart = nrest;
  qic.parity = EVENANDODD; qic.start_flag = 1; qic.nsrc = nsrc;
  qic.resid = resid; qic.relresid = 0; qic.deflate = 0;
  for (k = 0; k < nsrc; k++) {
    s[k] = (su3_vector *)src + (size_t)k * sites_on_node;
    d[k] = (su3_vector *)dst + (size_t)k * sites_on_node;
  }
  if (nsrc == 1) it = mat_invert_uml_field(s[0], d[0], &qic, (Real)m, h_fn);
  else it = mat_invert_block_uml(s, d, (Real)m, nsrc, &qic, h_fn);
  h_unpack(&qic, out);
  free(s); free(d);
  return it;
}

/* The same with low-mode deflation: eigvec = nvecs contiguous fields of
 * sites_on_node su3_vectors (both parities), eigval their eigenvalues of -D_eo D_oe; qic.deflate = 1 makes
 * mat_invert_uml_field start each CG from deflate()'s trial solution (mat_invert.c:131-183,341-383). */
int milcref_mat_invert_uml_deflated(const Real *src, Real *dst, double m, int max, int nrest, double resid,
                                    int nvecs, const Real *eigvec, const double *eigval, double *out) {
  quark_invert_control qic;
  int k, it;
  memset(&qic, 0, sizeof(qic));
  qic.prec = MILC_PRECISION; qic.min = 0; qic.max = max; qic.nrestart = nrest;
  qic.parity = EVENANDODD; qic.start_flag = 1; qic.nsrc = 1;
  qic.resid = resid; qic.relresid = 0; qic.deflate = 1;
  eigVec = (su3_vector **)malloc(nvecs * sizeof(su3_vector *));
  eigVal = (double *)malloc(nvecs * sizeof(double));
  for (k = 0; k < nvecs; k++) {
    eigVec[k] = (su3_vector *)eigvec + (size_t)k * sites_on_node;
    eigVal[k] = eigval[k];
  }
  param.eigen_param.Nvecs = nvecs;
  it = mat_invert_uml_field((su3_vector *)src, (su3_vector *)dst, &qic, (Real)m, h_fn);
  h_unpack(&qic, out);
  param.eigen_param.Nvecs = 0;
  free(eigVec); free(eigVal);
  eigVec = NULL; eigVal = NULL;
  return it;
}

/* ---- HISQ fermion force ---------------------------------------------
 * eo_fermion_force_multi, generic_ks/fermion_force_hisq_multi.c:170-216 (the wrapper_mx path of the
 * RHMC build: outer products of the nterms solution vectors, level-2 smearing force, derivative of
 * the U(3) projection, level-1 smearing force, projection onto the momenta):
 *     mom_mu(x) += eps * (traceless antihermitian part of the force)
 * links: thin links su3_matrix[4*V] with phases in; multi_x: nterms contiguous fields of su3_vector
 * (both parities: even = solution, odd = D solution, as update_h_rhmc.c:75-86 prepares them);
 * mom: anti_hermitmat[4*V] as 10 Reals each {m01.re,m01.im,m02.re,m02.im,m12.re,m12.im,m00im,m11im,
 * m22im,space} (include/su3.h), zero on entry here, the update on exit.
 * Returns the number of links whose force took the SVD / filter branches (sum). */
int milcref_hisq_force_naik(const Real *links, const Real *multi_x, const Real *residues, int n_naiks,
                            const int *n_orders, const double *eps_naik_in, double eps, Real *mom, Real *fat_lng);

int milcref_hisq_force(const Real *links, const Real *multi_x, const Real *residues, int nterms, double eps,
                       Real *mom) {
  double eps_naik[1] = {0.0};
  return milcref_hisq_force_naik(links, multi_x, residues, 1, &nterms, eps_naik, eps, mom, NULL);
}

/* The same with several Naik epsilons (the charm quark's mass-dependent correction): the terms come in
 * n_naiks classes of n_orders[k] terms, class k solved with the links of eps_naik[k] (eps_naik[0] = 0);
 * fermion_force_hisq_multi.c:1285-1375.  fat_lng, if not NULL: [n_naiks][2][4*V] su3_matrix, the fat and
 * long links of every class (load_hisq_fn_links, fermion_links_hisq_load_milc.c:573-636). */
int milcref_hisq_force_naik(const Real *links, const Real *multi_x, const Real *residues, int n_naiks,
                            const int *n_orders, const double *eps_naik_in, double eps, Real *mom, Real *fat_lng) {
  double eps_naik[MAX_NAIK];
  fermion_links_t *fl;
  su3_vector **xx;
  Real *res;
  size_t i;
  int k, dir, nterms = 0;
  if (!h_ready || n_naiks < 1 || n_naiks > MAX_NAIK) return -1;
  for (k = 0; k < n_naiks; k++) {
    eps_naik[k] = eps_naik_in[k];
    n_orders_naik[k] = n_orders[k];
    nterms += n_orders[k];
  }
  xx = (su3_vector **)malloc(nterms * sizeof(*xx));
  res = (Real *)malloc(nterms * sizeof(Real));
  for (k = 0; k < nterms; k++) {
    xx[k] = (su3_vector *)multi_x + (size_t)k * sites_on_node;
    res[k] = residues[k];
  }
  n_order_naik_total = nterms;
  hisq_svd_counter = 0;
  hisq_force_filter_counter = 0;
  for (i = 0; i < sites_on_node; i++) {
    memcpy(lattice[i].link, (const su3_matrix *)links + 4 * i, 4 * sizeof(su3_matrix));
    memset(lattice[i].mom, 0, 4 * sizeof(anti_hermitmat));
  }
  fl = create_fermion_links_hisq(MILC_PRECISION, n_naiks, eps_naik, phases_in, (su3_matrix *)links);
  if (fat_lng) {
    const size_t bytes = sizeof(su3_matrix) * 4 * sites_on_node;
    for (k = 0; k < n_naiks; k++) {
      fn_links_t *f = get_fm_links(fl)[k];
      memcpy((char *)fat_lng + (2 * k) * bytes, get_fatlinks(f), bytes);
      memcpy((char *)fat_lng + (2 * k + 1) * bytes, get_lnglinks(f), bytes);
    }
  }
  eo_fermion_force_multi((Real)eps, res, xx, nterms, MILC_PRECISION, fl);
  for (i = 0; i < sites_on_node; i++)
    for (dir = 0; dir < 4; dir++) memcpy(mom + 10 * (4 * i + dir), &lattice[i].mom[dir], sizeof(anti_hermitmat));
  destroy_fermion_links_hisq(fl);
  free(xx); free(res);
  n_orders_naik[0] = 0;
  return hisq_svd_counter + hisq_force_filter_counter;
}
