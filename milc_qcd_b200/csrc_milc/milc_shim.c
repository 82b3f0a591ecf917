/* milc_shim.c -- route-1 boundary: MILC's GPU solver symbols on the b200ks C ABI.
 *
 * Same job as the reference's QUDA glue (generic_ks/d_congrad5_fn_gpu.c:35-172,
 * generic_ks/ks_multicg_offset_gpu.c:26-252, generic_ks/dslash_fn.c:306-344): initialise the
 * qic outputs, take the zero-source shortcut on the host, keep the device link cache in
 * step with `fn`, call the solver, write back the results.  Differences by design: the
 * solver behind it follows the CPU algorithm (restarts, true-residual stop, iteration
 * counting), so qic->final_restart, size_r and total_iters are filled like the CPU path does.
 *
 * Standalone: cc -Iinclude milc_shim.c -lb200ks  (mirror types from include/b200ks_milc.h).
 * In a MILC tree: compile with -DB200KS_IN_MILC next to generic_ks_includes.h.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef B200KS_IN_MILC
#include "generic_ks_includes.h"
#include "b200ks_milc.h"
#define NX nx
#define NY ny
#define NZ nz
#define NT nt
#define SITES ((size_t)sites_on_node)
#define TOTAL_ITERS total_iters
#define FATAL(status) terminate(status)
#if defined(MAX_MIXED)
#define MIXED 2
#elif defined(HALF_MIXED)
#define MIXED 1
#else
#define MIXED 0
#endif
#else
#include "b200ks_milc.h"
static int s_dims[4] = {0, 0, 0, 0};
static int s_total_iters = 0;
static int s_mixed = 0;
#define NX s_dims[0]
#define NY s_dims[1]
#define NZ s_dims[2]
#define NT s_dims[3]
#define SITES ((size_t)s_dims[0] * s_dims[1] * s_dims[2] * s_dims[3])
#define TOTAL_ITERS s_total_iters
#define MIXED s_mixed
static void FATAL(int status) {
  printf("Termination: node 0, status = %d\n", status);
  fflush(stdout);
  exit(status);
}
#endif

#include "b200ks.h"

static b200ks_ctx *s_ctx = NULL;
static imp_ferm_links_t *fn_last = NULL; /* ks_multicg_offset_gpu.c:26-36 */

imp_ferm_links_t *get_fn_last(void) { return fn_last; }
void set_fn_last(imp_ferm_links_t *fn_last_new) { fn_last = fn_last_new; }

#ifndef B200KS_IN_MILC
void b200ks_milc_setup(int lx, int ly, int lz, int lt, int mixed_precision) {
  if (s_ctx && (lx != NX || ly != NY || lz != NZ || lt != NT)) {
    b200ks_destroy(s_ctx);
    s_ctx = NULL;
    fn_last = NULL;
  }
  s_dims[0] = lx; s_dims[1] = ly; s_dims[2] = lz; s_dims[3] = lt;
  s_mixed = mixed_precision;
}
void b200ks_milc_finalize(void) {
  if (s_ctx) b200ks_destroy(s_ctx);
  s_ctx = NULL;
  fn_last = NULL;
}
int b200ks_milc_total_iters(void) { return s_total_iters; }
b200ks_ctx *b200ks_milc_context(void) { return s_ctx; } /* diagnostics: b200ks_call_profile, b200ks_launch_count, ... */
#endif

static void die(const char *myname) {
  printf("%s(0): libb200ks: %s\n", myname, b200ks_last_error());
  FATAL(1);
}

static b200ks_ctx *context(const char *myname) {
  if (s_ctx == NULL) {
    int dims[4];
    dims[0] = NX; dims[1] = NY; dims[2] = NZ; dims[3] = NT;
    /* B200KS_DEVICE: first device; B200KS_NGPU=N (read by b200ks_create): spread the lattice over N devices */
    s_ctx = b200ks_create(dims, getenv("B200KS_DEVICE") ? atoi(getenv("B200KS_DEVICE")) : 0);
    if (s_ctx == NULL) die(myname);
  }
  return s_ctx;
}

/* d_congrad5_fn_gpu.c:121-126: refresh the device links when fn changed or was rebuilt -- and
 * also when the arrays were edited in place without notice (boundary_twist_fn,
 * fermion_links_fn_twist_milc.c:318-400).  b200ks_links_sync checks the content fingerprints on host
 * threads while the solve runs (B200KS_ALWAYS_RELOAD_LINKS: 0 trust MILC's flag, 1 upload always,
 * 2 = default, 3 check before the solve). */
static void refresh_links(const char *myname, imp_ferm_links_t *fn) {
  static int mode = -1;
  if (mode < 0) {
    const char *e = getenv("B200KS_ALWAYS_RELOAD_LINKS");
    mode = e ? atoi(e) : 2;
    if (mode < 0 || mode > 3) mode = 2;
  }
  if (b200ks_links_sync(context(myname), fn->fat, fn->lng, MILC_PRECISION, fn != fn_last || fn->notify_quda_new_links, mode) < 0)
    die(myname);
  fn->notify_quda_new_links = 0; /* cancel_quda_notification(fn) */
  fn_last = fn;
}

/* the reference's trivial-solution test (d_congrad5_fn_gpu.c:63-89) only needs to know whether the
 * source vanishes: stop at the first nonzero number (a 50 MB norm on one host thread is 10 % of a solve) */
static int source_is_zero(const su3_vector *v, int parity) {
  size_t lo = (parity == ODD) ? SITES / 2 : 0, hi = (parity == EVEN) ? SITES / 2 : SITES, i;
  const Real *r = (const Real *)v;
  for (i = 6 * lo; i < 6 * hi; i++)
    if (r[i] != 0) return 0;
  return 1;
}

int ks_congrad_parity_gpu(su3_vector *t_src, su3_vector *t_dest, quark_invert_control *qic, Real mass,
                          imp_ferm_links_t *fn) {
  char myname[] = "ks_congrad_parity_gpu";
  b200ks_invert_args a;
  b200ks_invert_result r;
  int iters;

  qic->size_r = 0;
  qic->size_relr = 0;
  qic->final_iters = 0;
  qic->final_restart = 0;
  qic->converged = 1;
  qic->final_rsq = 0.;
  qic->final_relrsq = 0.;

  if (fn == NULL) {
    printf("%s(0): Called with NULL fn\n", myname);
    FATAL(1);
  }
  if (qic->parity != EVEN && qic->parity != ODD) {
    printf("%s: Unrecognised parity\n", myname);
    FATAL(2);
  }
  /* trivial solution (d_congrad5_fn_gpu.c:63-89) */
  if (source_is_zero(t_src, qic->parity)) {
    size_t lo = (qic->parity == ODD) ? SITES / 2 : 0;
    memset(t_dest + lo, 0, (SITES / 2) * sizeof(su3_vector));
    return 0;
  }
  refresh_links(myname, fn);
  memset(&a, 0, sizeof(a));
  a.parity = qic->parity;
  a.max_iter = qic->max;
  a.nrestart = qic->nrestart;
  a.resid = qic->resid;
  a.relresid = qic->relresid;
  a.mixed_precision = (qic->prec == 1) ? (MIXED ? MIXED : 1) : MIXED;
  if (MILC_PRECISION == 2 && qic->prec == 2) a.mixed_precision = MIXED;
  iters = b200ks_congrad(context(myname), t_src, t_dest, (double)mass, &a, &r, MILC_PRECISION);
  if (iters < 0) die(myname);
  qic->final_rsq = (Real)r.final_rsq;
  qic->final_relrsq = (Real)r.final_relrsq;
  qic->size_r = (Real)r.size_r;
  qic->size_relr = (Real)r.size_relr;
  qic->final_iters = r.final_iters;
  qic->final_restart = r.final_restart;
  qic->converged = r.converged;
  TOTAL_ITERS += iters;
  return iters;
}

/* Block solver: all sources go to the multi-right-hand-side CG (b200ks_congrad_block).  The CPU
 * reference is a loop over sources (d_congrad5_fn_milc.c:409-417); qic reports what the loop
 * would leave behind where that is well defined (total iterations, converged = all converged)
 * and the worst residual, like the QUDA glue (d_congrad5_fn_gpu.c:283-296). */
#define B200KS_MAX_BLOCK 64
int ks_congrad_block_parity_gpu(int nsrc, su3_vector **t_src, su3_vector **t_dest, quark_invert_control *qic,
                                Real mass, imp_ferm_links_t *fn) {
  char myname[] = "ks_congrad_block_parity_gpu";
  b200ks_invert_args a;
  b200ks_invert_result r[B200KS_MAX_BLOCK];
  int iters, k;

  qic->size_r = 0;
  qic->size_relr = 0;
  qic->final_iters = 0;
  qic->final_restart = 0;
  qic->converged = 1;
  qic->final_rsq = 0.;
  qic->final_relrsq = 0.;
  if (nsrc <= 0) return 0;
  if (nsrc > B200KS_MAX_BLOCK) { /* very wide blocks: in chunks, qic = worst of all chunks */
    int done = 0, tot = 0;
    quark_invert_control acc = *qic, q;
    while (done < nsrc) {
      int n = nsrc - done < B200KS_MAX_BLOCK ? nsrc - done : B200KS_MAX_BLOCK;
      q = *qic;
      tot += ks_congrad_block_parity_gpu(n, t_src + done, t_dest + done, &q, mass, fn);
      if (q.final_rsq > acc.final_rsq) acc.final_rsq = q.final_rsq;
      if (q.final_relrsq > acc.final_relrsq) acc.final_relrsq = q.final_relrsq;
      if (q.size_r > acc.size_r) acc.size_r = q.size_r;
      if (q.size_relr > acc.size_relr) acc.size_relr = q.size_relr;
      if (q.final_restart > acc.final_restart) acc.final_restart = q.final_restart;
      if (!q.converged) acc.converged = 0;
      done += n;
    }
    acc.final_iters = tot;
    *qic = acc;
    return tot;
  }
  if (fn == NULL) {
    printf("%s(0): Called with NULL fn\n", myname);
    FATAL(1);
  }
  if (qic->parity != EVEN && qic->parity != ODD) {
    printf("%s: Unrecognised parity\n", myname);
    FATAL(2);
  }
  refresh_links(myname, fn);
  memset(&a, 0, sizeof(a));
  a.parity = qic->parity;
  a.max_iter = qic->max;
  a.nrestart = qic->nrestart;
  a.resid = qic->resid;
  a.relresid = qic->relresid;
  a.mixed_precision = (qic->prec == 1) ? (MIXED ? MIXED : 1) : MIXED;
  if (MILC_PRECISION == 2 && qic->prec == 2) a.mixed_precision = MIXED;
  iters = b200ks_congrad_block(context(myname), nsrc, (const void *const *)t_src, (void *const *)t_dest, (double)mass, &a, r,
                               MILC_PRECISION);
  if (iters < 0) die(myname);
  for (k = 0; k < nsrc; k++) {
    if (r[k].final_rsq > qic->final_rsq) qic->final_rsq = (Real)r[k].final_rsq;
    if (r[k].final_relrsq > qic->final_relrsq) qic->final_relrsq = (Real)r[k].final_relrsq;
    if (r[k].size_r > qic->size_r) qic->size_r = (Real)r[k].size_r;
    if (r[k].size_relr > qic->size_relr) qic->size_relr = (Real)r[k].size_relr;
    if (r[k].final_restart > qic->final_restart) qic->final_restart = r[k].final_restart;
    if (!r[k].converged) qic->converged = 0;
  }
  qic->final_iters = iters;
  TOTAL_ITERS += iters;
  return iters;
}

/* Low modes for deflation.  MILC keeps them in application globals (eigVec, eigVal,
 * param.eigen_param.Nvecs: ks_spectrum/lattice.h, params.h) that a library cannot name, so the application hands
 * them over once after reading or computing them; they are uploaded and stay in HBM until replaced.
 * eigvec[j]: su3_vector[sites_on_node], both parities filled; eigval[j]: eigenvalue of -D_eo D_oe.  nvecs = 0 drops
 * them.  The UML sequences below then honour qic->deflate like mat_invert.c:341-353,376-387,428-437 do. */
static int s_neig = 0;
static int *s_eig_handle = NULL;
void b200ks_milc_set_eigenvectors(int nvecs, su3_vector **eigvec, double *eigval) {
  char myname[] = "b200ks_milc_set_eigenvectors";
  b200ks_ctx *ctx = context(myname);
  int j;
  if (b200ks_eig_set(ctx, 0, NULL, NULL, 0) < 0) die(myname);
  for (j = 0; j < s_neig; j++)
    if (b200ks_vec_free(ctx, s_eig_handle[j]) < 0) die(myname);
  free(s_eig_handle);
  s_eig_handle = NULL;
  s_neig = 0;
  if (nvecs <= 0) return;
  s_eig_handle = (int *)malloc(nvecs * sizeof(int));
  for (j = 0; j < nvecs; j++) {
    s_eig_handle[j] = b200ks_vec_create(ctx);
    if (s_eig_handle[j] < 0) die(myname);
    s_neig = j + 1;
    if (b200ks_vec_upload(ctx, s_eig_handle[j], eigvec[j], EVENANDODD, MILC_PRECISION) < 0) die(myname);
  }
  if (b200ks_eig_set(ctx, nvecs, s_eig_handle, eigval, 0) < 0) die(myname);
}

/* mat_invert_uml_field / mat_invert_block_uml (generic_ks/mat_invert.c:328-402,409-475) as one
 * device-resident sequence: M^+ src, [deflation,] even solve, odd reconstruction, [deflation,] odd polish. */
int mat_invert_block_uml_gpu(su3_vector **src, su3_vector **dst, Real mass, int nsrc, quark_invert_control *qic,
                             imp_ferm_links_t *fn) {
  char myname[] = "mat_invert_block_uml_gpu";
  b200ks_invert_args a;
  b200ks_invert_result r[2 * B200KS_MAX_BLOCK];
  int iters, k;

  qic->size_r = 0;
  qic->size_relr = 0;
  qic->final_iters = 0;
  qic->final_restart = 0;
  qic->converged = 1;
  qic->final_rsq = 0.;
  qic->final_relrsq = 0.;
  if (nsrc <= 0) return 0;
  if (nsrc > B200KS_MAX_BLOCK) {
    printf("%s: more than %d sources\n", myname, B200KS_MAX_BLOCK);
    FATAL(1);
  }
  if (fn == NULL) {
    printf("%s(0): Called with NULL fn\n", myname);
    FATAL(1);
  }
  refresh_links(myname, fn);
  memset(&a, 0, sizeof(a));
  a.parity = EVEN;
  a.max_iter = qic->max;
  a.nrestart = qic->nrestart;
  a.resid = qic->resid;
  a.relresid = qic->relresid;
  a.mixed_precision = (qic->prec == 1) ? (MIXED ? MIXED : 1) : MIXED;
  if (MILC_PRECISION == 2 && qic->prec == 2) a.mixed_precision = MIXED;
  if (b200ks_eig_use_in_uml(context(myname), qic->deflate && s_neig > 0) < 0) die(myname);
  iters = b200ks_mat_invert_uml(context(myname), nsrc, (const void *const *)src, (void *const *)dst, (double)mass, &a, r,
                                MILC_PRECISION);
  if (iters < 0) die(myname);
  for (k = 0; k < 2 * nsrc; k++) {
    if (r[k].final_rsq > qic->final_rsq) qic->final_rsq = (Real)r[k].final_rsq;
    if (r[k].final_relrsq > qic->final_relrsq) qic->final_relrsq = (Real)r[k].final_relrsq;
    if (r[k].size_r > qic->size_r) qic->size_r = (Real)r[k].size_r;
    if (r[k].size_relr > qic->size_relr) qic->size_relr = (Real)r[k].size_relr;
    if (r[k].final_restart > qic->final_restart) qic->final_restart = r[k].final_restart;
    if (!r[k].converged) qic->converged = 0;
  }
  qic->final_iters = iters;
  qic->parity = ODD; /* the state the reference leaves behind (mat_invert.c:392) */
  TOTAL_ITERS += iters;
  return iters;
}

int mat_invert_uml_field_gpu(su3_vector *src, su3_vector *dst, quark_invert_control *qic, Real mass,
                             imp_ferm_links_t *fn) {
  return mat_invert_block_uml_gpu(&src, &dst, mass, 1, qic, fn);
}

/* Incremental eigCG (generic_ks/inc_eigcg.c:851-950): one solve of the sequence on the device, then eigVec[],
 * eigVal[] (untouched, as in the reference, until calc_eigenpairs) and eigcgp are brought up to date on the host. */
#ifdef B200KS_IN_MILC
typedef double_complex b200ks_double_complex;
#endif
static int s_eigcg_m = 0, s_eigcg_nvecs = 0, s_eigcg_max = 0;
int ks_inc_eigCG_parity_gpu(su3_vector *src, su3_vector *dest, double *eigVal, su3_vector **eigVec, eigcg_params *eigcgp,
                            quark_invert_control *qic, Real mass, imp_ferm_links_t *fn) {
  char myname[] = "ks_inc_eigCG_parity_gpu";
  b200ks_invert_args a;
  b200ks_invert_result r;
  b200ks_ctx *ctx;
  int iters, j, k, n_old, n_new, ld;
  double *Hdev;
  (void)eigVal;

  qic->size_r = 0;
  qic->size_relr = 1.;
  qic->final_iters = 0;
  qic->final_restart = 0;
  qic->converged = 1;
  qic->final_rsq = 0.;
  qic->final_relrsq = 0.;
  if (fn == NULL) {
    printf("%s(0): Called with NULL fn\n", myname);
    FATAL(1);
  }
  if (qic->parity != EVEN && qic->parity != ODD) {
    printf("%s: Unrecognised parity\n", myname);
    FATAL(2);
  }
  ctx = context(myname);
  refresh_links(myname, fn);
  n_old = eigcgp->Nvecs_curr;
  ld = eigcgp->Nvecs_max;
  if (n_old == 0 || eigcgp->m != s_eigcg_m || ld != s_eigcg_max || b200ks_eigcg_count(ctx) != n_old) {
    /* a new sequence (inc_eigcg.c:868-873 allocates H here) */
    if (b200ks_eigcg_init(ctx, eigcgp->m, eigcgp->Nvecs, ld) < 0) die(myname);
    s_eigcg_m = eigcgp->m; s_eigcg_nvecs = eigcgp->Nvecs; s_eigcg_max = ld;
    if (eigcgp->H != NULL) free(eigcgp->H);
    eigcgp->H = (b200ks_double_complex *)calloc((size_t)ld * ld, sizeof(b200ks_double_complex));
    n_old = 0;
  }
  memset(&a, 0, sizeof(a));
  a.parity = qic->parity;
  a.max_iter = qic->max;
  a.nrestart = qic->nrestart;
  a.resid = qic->resid;
  a.relresid = qic->relresid;
  iters = b200ks_inc_eigcg(ctx, src, dest, (double)mass, &a, &r, MILC_PRECISION);
  if (iters < 0) die(myname);
  qic->final_rsq = (Real)r.final_rsq;
  qic->final_relrsq = (Real)r.final_relrsq;
  qic->size_r = (Real)r.size_r;
  qic->size_relr = (Real)r.size_relr;
  qic->final_iters = r.final_iters;
  qic->final_restart = r.final_restart;
  qic->converged = r.converged;
  TOTAL_ITERS += iters;
  n_new = b200ks_eigcg_count(ctx);
  for (j = n_old; j < n_new; j++)
    if (b200ks_eigcg_vec_download(ctx, j, eigVec[j], MILC_PRECISION) < 0) die(myname);
  Hdev = (double *)malloc(sizeof(double) * 2 * (size_t)ld * ld);
  if (b200ks_eigcg_hmatrix(ctx, Hdev) < 0) die(myname);
  for (j = 0; j < n_new; j++)      /* row-major [k][j] -> MILC's column-major H[k + ld*j] */
    for (k = 0; k < n_new; k++) {
      eigcgp->H[k + (size_t)ld * j].real = Hdev[2 * ((size_t)k * ld + j)];
      eigcgp->H[k + (size_t)ld * j].imag = Hdev[2 * ((size_t)k * ld + j) + 1];
    }
  free(Hdev);
  eigcgp->Nvecs_curr = n_new;
  eigcgp->Nvecs = (ld - n_new < eigcgp->Nvecs) ? (ld - n_new) : eigcgp->Nvecs;
  return iters;
}

/* calc_eigenpairs (inc_eigcg.c:282-300): Rayleigh-Ritz on the accumulated vectors; eigVal[], eigVec[] and H follow */
void calc_eigenpairs_gpu(double *eigVal, su3_vector **eigVec, eigcg_params *eigcgp, int parity) {
  char myname[] = "calc_eigenpairs_gpu";
  b200ks_ctx *ctx = context(myname);
  int j, k, n, ld = eigcgp->Nvecs_max;
  (void)parity;
  n = b200ks_eigcg_pairs(ctx, eigVal, eigcgp->Nvecs_curr);
  if (n < 0) die(myname);
  for (j = 0; j < n; j++) {
    if (b200ks_eigcg_vec_download(ctx, j, eigVec[j], MILC_PRECISION) < 0) die(myname);
    for (k = 0; k < j; k++) eigcgp->H[k + (size_t)ld * j].real = eigcgp->H[k + (size_t)ld * j].imag = 0.;
    eigcgp->H[j + (size_t)ld * j].real = eigVal[j];
    eigcgp->H[j + (size_t)ld * j].imag = 0.;
  }
}

/* ---- meson tie-ups (generic_ks/ks_meson_mom.c:160-437) ------------------------------------------------------ */
/* gamma bits of a LOCAL sink operator, -1 when the operator needs link shifts.  Compatibility indices: enum
 * spin_taste_type (generic_ks/spin_taste_ops.c:805-853; pion5 = 0, pion05 = 1, rhoi = 8 ... rhoz0 = 15) with the
 * operators spin_taste_op_links gives them (:1205-1270); gamma-gamma indices (>= 128, :960-973) are local when spin
 * and taste agree; bits = gamma_hex_value (generic_wilson/gammas.c:21-22). */
static int local_spin_bits(int index) {
  static const int hex[16] = {1, 2, 4, 8, 15, 6, 5, 3, 9, 10, 12, 14, 13, 11, 7, 0};
  if (index >= 128) {
    int s = (index - 128) / 16, t = (index - 128) % 16;
    return (s == t && s < 16) ? hex[s] : -1;
  }
  switch (index) {
  case 0: return 15;            /* pion5: gamma_5 x gamma_5 */
  case 1: return 0;             /* pion05: 1 x 1 */
  case 9: return 1;             /* rhox */
  case 10: return 2;            /* rhoy */
  case 8: case 11: return 4;    /* rhoi, rhoz */
  case 13: return 9;            /* rhox0: gamma_x gamma_t */
  case 14: return 10;           /* rhoy0 */
  case 12: case 15: return 12;  /* rhoi0, rhoz0 */
  default: return -1;
  }
}
/* spin_taste_ops.c:983-1035: rhoxsfn .. rhotsfn = 22..25, ffn 26..29, bfn 30..33, ape 34..37, fape 38..41, bape 42..45 */
static int st_is_rhosfn(int i) { return (i >= 22 && i <= 25) || (i >= 34 && i <= 37); }
static int st_is_rhosffn(int i) { return (i >= 26 && i <= 29) || (i >= 38 && i <= 41); }
static int st_is_rhosbfn(int i) { return (i >= 30 && i <= 33) || (i >= 42 && i <= 45); }
static int st_forward(int i) { return st_is_rhosfn(i) ? i + 4 : -1; }    /* forward_index(), :1041-1064 */
static int st_backward(int i) { return st_is_rhosfn(i) ? i + 8 : -1; }   /* backward_index(), :1067-1090 */

#ifdef B200KS_IN_MILC
#define SINK_OP(fn, index, r0, dest, src) spin_taste_op_fn(fn, index, r0, dest, src)
#else
static void SINK_OP(imp_ferm_links_t *fn, int index, int r0[], su3_vector *dest, su3_vector *src) {
  (void)fn; (void)r0; (void)dest; (void)src;
  printf("ks_meson_cont_mom_gpu: sink operator %d needs MILC's spin_taste_op_fn (build inside the MILC tree)\n", index);
  FATAL(1);
}
#endif

void ks_meson_cont_mom_gpu(B200KS_MILC_COMPLEX **prop, su3_vector *src1, su3_vector *src2, int no_q_momenta, int **q_momstore,
                           char **q_parity, int no_spin_taste_corr, int num_corr_mom[], int **corr_table, int p_index[],
                           imp_ferm_links_t *fn_src1, imp_ferm_links_t *fn_src2, int spin_taste_snk[], int meson_phase[],
                           Real meson_factor[], int corr_index[], int r0[]) {
  char myname[] = "meson_cont_mom";
  b200ks_ctx *ctx = context(myname);
  const int nt_ = NT;
  int g, k, t, d;
  int *mom;
  char *par;
  double *corr, *corr2 = NULL;
  su3_vector *work = NULL;

  if (no_q_momenta > 100) { /* MAXQ, ks_meson_mom.c:225-230 */
    printf("%s(%d): no_q_momenta %d exceeds max %d\n", myname, 0, no_q_momenta, 100);
    FATAL(1);
  }
  mom = (int *)malloc(3 * sizeof(int) * (size_t)(no_q_momenta > 0 ? no_q_momenta : 1));
  par = (char *)malloc(3 * (size_t)(no_q_momenta > 0 ? no_q_momenta : 1));
  corr = (double *)malloc(2 * sizeof(double) * (size_t)nt_ * (no_q_momenta > 0 ? no_q_momenta : 1));
  if (mom == NULL || par == NULL || corr == NULL) {
    printf("%s(%d): No room for meson\n", myname, 0);
    FATAL(1);
  }
  for (g = 0; g < no_spin_taste_corr; g++) {
    const int nk = num_corr_mom[g];
    const int st = spin_taste_snk[corr_table[g][0]];
    const int bits = local_spin_bits(st);
    int rc;
    if (nk < 1) continue;
    for (k = 0; k < nk; k++) {
      const int p = p_index[corr_table[g][k]];
      for (d = 0; d < 3; d++) { mom[3 * k + d] = q_momstore[p][d]; par[3 * k + d] = q_parity[p][d]; }
    }
    if (bits >= 0) {
      rc = b200ks_meson_mom(ctx, src1, src2, MILC_PRECISION, bits, r0, nk, mom, par, corr);
    } else {
      if (work == NULL) work = (su3_vector *)malloc(SITES * sizeof(su3_vector));
      if (work == NULL) { printf("%s(%d): No room for meson\n", myname, 0); FATAL(1); }
      if (st_is_rhosfn(st)) { /* (db + df)/2, ks_meson_mom.c:296-312, 349-356 */
        if (corr2 == NULL) corr2 = (double *)malloc(2 * sizeof(double) * (size_t)nt_ * no_q_momenta);
        SINK_OP(fn_src1, st_backward(st), r0, work, src1);
        rc = b200ks_meson_mom(ctx, work, src2, MILC_PRECISION, -1, r0, nk, mom, par, corr);
        if (rc >= 0) {
          SINK_OP(fn_src2, st_forward(st), r0, work, src2);
          rc = b200ks_meson_mom(ctx, src1, work, MILC_PRECISION, -1, r0, nk, mom, par, corr2);
          for (k = 0; k < 2 * nt_ * nk; k++) corr[k] = 0.5 * (corr[k] + corr2[k]);
        }
      } else if (st_is_rhosffn(st)) {
        SINK_OP(fn_src2, st_forward(st), r0, work, src2);
        rc = b200ks_meson_mom(ctx, src1, work, MILC_PRECISION, -1, r0, nk, mom, par, corr);
      } else {
        SINK_OP(fn_src1, st_is_rhosbfn(st) ? st_backward(st) : st, r0, work, src1);
        rc = b200ks_meson_mom(ctx, work, src2, MILC_PRECISION, -1, r0, nk, mom, par, corr);
      }
    }
    if (rc < 0) die(myname);
    /* norm_v (ks_meson_mom.c:100-131) and the accumulation behind it (:405-419) */
    for (t = 0; t < nt_; t++)
      for (k = 0; k < nk; k++) {
        const int c = corr_table[g][k];
        const double re = corr[2 * ((size_t)t * nk + k)], im = corr[2 * ((size_t)t * nk + k) + 1];
        const Real fact = meson_factor[c];
        double zr, zi;
        switch (meson_phase[c]) {
        case 0: zr = re; zi = im; break;
        case 1: zr = -im; zi = re; break;     /* TIMESPLUSI */
        case 2: zr = -re; zi = -im; break;    /* TIMESMINUSONE */
        default: zr = im; zi = -re; break;    /* TIMESMINUSI */
        }
        prop[corr_index[c]][t].real += (Real)(zr * fact);
        prop[corr_index[c]][t].imag += (Real)(zi * fact);
      }
  }
  free(mom); free(par); free(corr); free(corr2); free(work);
}

int ks_multicg_offset_field_gpu(su3_vector *src, su3_vector **psim, ks_param *ksp, int num_offsets,
                                quark_invert_control *qic, imp_ferm_links_t *fn) {
  char myname[] = "ks_multicg_offset_field_gpu";
  b200ks_invert_args a;
  b200ks_invert_result r[B200KS_MAX_SHIFTS];
  double offsets[B200KS_MAX_SHIFTS];
  int j, iters;

  if (qic[0].relresid != 0.) {
    printf("%s: GPU code does not yet support a Fermilab-type relative residual\n", myname);
    FATAL(1);
  }
  for (j = 0; j < num_offsets; j++) {
    qic[j].final_rsq = 0.;
    qic[j].final_relrsq = 0.;
    qic[j].size_r = 0.;
    qic[j].size_relr = 0.;
    qic[j].final_iters = 0;
    qic[j].final_restart = 0;
    qic[j].converged = 1;
  }
  if (num_offsets == 0) return 0;
  if (num_offsets > B200KS_MAX_SHIFTS) {
    printf("%s: more than %d offsets\n", myname, B200KS_MAX_SHIFTS);
    FATAL(1);
  }
  if (fn == NULL) {
    printf("%s(0): Called with NULL fn\n", myname);
    FATAL(1);
  }
  if (qic[0].parity == EVENANDODD) {
    printf("%s: EVENANDODD not supported\n", myname);
    FATAL(1);
  }
  if (qic[0].parity != EVEN && qic[0].parity != ODD) {
    printf("%s: Unrecognised parity\n", myname);
    FATAL(2);
  }
  for (j = 0; j < num_offsets; j++) {
    if (ksp[j].offset <= 0) { /* ks_multicg_offset.c:139-145 */
      printf("ks_multicg_offset_field(0): Called with nonpositive offset %e\n", (double)ksp[j].offset);
      FATAL(1);
    }
    offsets[j] = ksp[j].offset;
  }
  if (source_is_zero(src, qic[0].parity)) {
    size_t lo = (qic[0].parity == ODD) ? SITES / 2 : 0;
    for (j = 0; j < num_offsets; j++) memset(psim[j] + lo, 0, (SITES / 2) * sizeof(su3_vector));
    return 0;
  }
  refresh_links(myname, fn);
  memset(&a, 0, sizeof(a));
  a.parity = qic[0].parity;
  a.max_iter = qic[0].max;
  a.nrestart = qic[0].nrestart;
  a.resid = qic[0].resid;
  a.mixed_precision = MIXED ? 1 : 0; /* never half with multi-shift (ks_multicg_offset_gpu.c:165-170) */
  iters = b200ks_multicg(context(myname), src, (void *const *)psim, offsets, num_offsets, &a, r, MILC_PRECISION);
  if (iters < 0) die(myname);
  for (j = 0; j < num_offsets; j++) {
    qic[j].final_rsq = (Real)r[j].final_rsq;
    qic[j].size_r = (Real)r[j].size_r;
    qic[j].final_iters = r[j].final_iters;
    qic[j].converged = r[j].converged;
  }
  TOTAL_ITERS += iters;
  return iters;
}

void dslash_fn_field(su3_vector *src, su3_vector *dest, int parity, fn_links_t *fn) {
  char myname[] = "dslash_fn_field";
  if (fn == NULL) {
    printf("dslash_fn_field_special: invalid fn links!\n");
    FATAL(1);
  }
  if (parity != EVEN && parity != ODD && parity != EVENANDODD) {
    printf("%s: Unrecognised parity\n", myname);
    FATAL(2);
  }
  refresh_links(myname, fn);
  if (b200ks_dslash(context(myname), src, dest, parity, MILC_PRECISION) < 0) die(myname);
}
