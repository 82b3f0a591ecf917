/* oracle/ref_harness/eigcg_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * ctypes-friendly face of the reference's incremental eigCG (generic_ks/inc_eigcg.c, compiled unmodified
 * from /root/reference by oracle/build_ref.sh together with harness.c).  inc_eigcg.c needs LAPACK
 * (zheevx, zgeqrf, zungqr, zheev, zpotrf, zpotrs) and BLAS (zcopy, zhemm, zgemm): the build links the
 * OpenBLAS that ships inside the image's Python packages; without one this file is left out and the eigCG
 * oracle says "parity unpinned".
 *   milcref_eigcg      ks_eigCG_parity      (inc_eigcg.c:377-850)
 *   milcref_inc_eigcg  ks_inc_eigCG_parity  (inc_eigcg.c:851-950) with the eigcg_params state kept here
 *   milcref_eigcg_pairs calc_eigenpairs     (inc_eigcg.c:282-300)
 */
#include "generic_ks_includes.h"
#include <string.h>

extern fn_links_t *milcref_fn(void);

static eigcg_params h_p = {0, 0, 0, 0, NULL};
static su3_vector **h_vec = NULL;
static double *h_val = NULL;
static int h_nalloc = 0;

static void h_free(void) {
  int j;
  for (j = 0; j < h_nalloc; j++) free(h_vec[j]);
  free(h_vec); free(h_val);
  if (h_p.H) free(h_p.H);
  h_vec = NULL; h_val = NULL; h_nalloc = 0;
  h_p.H = NULL; h_p.Nvecs_curr = 0;
}

/* (re)start an incremental sequence: m search vectors, Nvecs pairs per solve, at most Nvecs_max in all */
int milcref_inc_eigcg_init(int m, int Nvecs, int Nvecs_max) {
  int j;
  h_free();
  h_p.m = m; h_p.Nvecs = Nvecs; h_p.Nvecs_curr = 0; h_p.Nvecs_max = Nvecs_max;
  h_nalloc = Nvecs_max + m;   /* ks_eigCG_parity writes m search vectors behind the current ones */
  h_vec = (su3_vector **)malloc(h_nalloc * sizeof(su3_vector *));
  h_val = (double *)calloc(h_nalloc, sizeof(double));
  for (j = 0; j < h_nalloc; j++) h_vec[j] = (su3_vector *)calloc(sites_on_node, sizeof(su3_vector));
  return 0;
}

static void h_qic(quark_invert_control *qic, int parity, int max, int nrest, double resid) {
  memset(qic, 0, sizeof(*qic));
  qic->prec = MILC_PRECISION; qic->max = max; qic->nrestart = nrest; qic->parity = parity;
  qic->start_flag = 1; qic->nsrc = 1; qic->resid = resid; qic->relresid = 0;
}
static void h_out(const quark_invert_control *qic, double *out) {
  out[0] = qic->final_rsq; out[1] = qic->final_relrsq; out[2] = qic->size_r; out[3] = qic->size_relr;
  out[4] = qic->final_iters; out[5] = qic->final_restart; out[6] = qic->converged;
}

/* one solve of the sequence; returns iterations; *ncurr = eigenvectors accumulated so far */
int milcref_inc_eigcg(const Real *src, Real *dest, double mass, int parity, int max, int nrest, double resid,
                      double *out, int *ncurr) {
  quark_invert_control qic;
  int it;
  h_qic(&qic, parity, max, nrest, resid);
  it = ks_inc_eigCG_parity((su3_vector *)src, (su3_vector *)dest, h_val, h_vec, &h_p, &qic, (Real)mass, milcref_fn());
  h_out(&qic, out);
  *ncurr = h_p.Nvecs_curr;
  return it;
}

/* Rayleigh-Ritz on everything accumulated: eigenvalues of -D^2 (ascending) and vectors [n][sites][3][2] */
int milcref_eigcg_pairs(int parity, double *eigval, Real *eigvec, double *H_out) {
  int j, n = h_p.Nvecs_curr;
  if (H_out) memcpy(H_out, h_p.H, sizeof(double) * 2 * h_p.Nvecs_max * h_p.Nvecs_max);
  calc_eigenpairs(h_val, h_vec, &h_p, parity);
  for (j = 0; j < n; j++) {
    eigval[j] = h_val[j];
    memcpy(eigvec + (size_t)j * 6 * sites_on_node, h_vec[j], sizeof(su3_vector) * sites_on_node);
  }
  return n;
}

/* the single-solve form: Nvecs lowest Ritz pairs of -D^2 from an m-vector search space */
int milcref_eigcg(const Real *src, Real *dest, double mass, int parity, int max, int nrest, double resid,
                  int m, int Nvecs, double *eigval, Real *eigvec, double *out) {
  quark_invert_control qic;
  int it, j;
  su3_vector **v = (su3_vector **)malloc(m * sizeof(su3_vector *));
  for (j = 0; j < m; j++) v[j] = (su3_vector *)calloc(sites_on_node, sizeof(su3_vector));
  h_qic(&qic, parity, max, nrest, resid);
  it = ks_eigCG_parity((su3_vector *)src, (su3_vector *)dest, eigval, v, m, Nvecs, &qic, (Real)mass, milcref_fn());
  h_out(&qic, out);
  for (j = 0; j < Nvecs; j++) memcpy(eigvec + (size_t)j * 6 * sites_on_node, v[j], sizeof(su3_vector) * sites_on_node);
  for (j = 0; j < m; j++) free(v[j]);
  free(v);
  return it;
}
