/* tests/host/b200ks_stub.c -- TEST INFRASTRUCTURE.  A recording stand-in for the part of the b200ks C ABI
 * (include/b200ks.h) that milc_qcd_b200/csrc_milc/milc_shim.c calls, so that tests/test_milc_shim_host.py can
 * exercise the shim's host logic (qic bookkeeping, zero-source shortcut, link-cache decisions, eigenvector
 * hand-over and the qic->deflate switch) without a GPU.  Nothing here computes physics: the "solvers" copy the
 * source into the solution on the requested parity and return canned iteration counts and residuals.
 * Built by the test:  gcc -shared -fPIC -Iinclude milc_shim.c tests/host/b200ks_stub.c -o libmilc_shim_stub.so */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200ks.h"

static char s_log[1 << 16];
static size_t s_len = 0;
static int s_dims[4];
static int s_next_vec = 0, s_live_vecs = 0;
static int s_fail_next = 0;

static void logf_(const char *fmt, ...);
#include <stdarg.h>
static void logf_(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  s_len += (size_t)vsnprintf(s_log + s_len, sizeof(s_log) - s_len, fmt, ap);
  va_end(ap);
  if (s_len >= sizeof(s_log)) s_len = sizeof(s_log) - 1;
}
const char *stub_log(void) { return s_log; }
void stub_reset(void) { s_len = 0; s_log[0] = 0; }
void stub_fail_next(int code) { s_fail_next = code; }
int stub_live_vecs(void) { return s_live_vecs; }

static size_t sites(void) { return (size_t)s_dims[0] * s_dims[1] * s_dims[2] * s_dims[3]; }

b200ks_ctx *b200ks_create(const int latsize[4], int device) {
  memcpy(s_dims, latsize, sizeof(s_dims));
  logf_("create %d %d %d %d dev %d\n", latsize[0], latsize[1], latsize[2], latsize[3], device);
  return (b200ks_ctx *)s_dims;
}
void b200ks_destroy(b200ks_ctx *c) { (void)c; logf_("destroy\n"); s_live_vecs = 0; }
int b200ks_num_gpus(b200ks_ctx *c) { (void)c; return 1; }
static void fill(b200ks_invert_result *r, int iters);
static void copy_parity(void *dst, const void *src, int parity, int host_prec);
static int s_eig_n = 0, s_eig_add = 0, s_eig_max = 0;
int b200ks_eigcg_init(b200ks_ctx *c, int m, int nvecs, int nmax) { (void)c; logf_("eigcg_init m %d nvecs %d max %d\n", m, nvecs, nmax); s_eig_n = 0; s_eig_add = nvecs; s_eig_max = nmax; return 0; }
int b200ks_eigcg_count(b200ks_ctx *c) { (void)c; return s_eig_n; }
int b200ks_inc_eigcg(b200ks_ctx *c, const void *src, void *dest, double mass, const b200ks_invert_args *a, b200ks_invert_result *r, int host_prec) {
  (void)c;
  logf_("inc_eigcg mass %g parity %d\n", mass, a->parity);
  copy_parity(dest, src, a->parity, host_prec);
  fill(r, 41);
  s_eig_n += s_eig_add;
  if (s_eig_n > s_eig_max) s_eig_n = s_eig_max;
  return 41;
}
int b200ks_eigcg_vec_download(b200ks_ctx *c, int j, void *host, int host_prec) {
  (void)c;
  logf_("eigcg_vec_download %d\n", j);
  if (host_prec == 2) ((double *)host)[0] = 100.0 + j; else ((float *)host)[0] = 100.0f + j;
  return 0;
}
int b200ks_eigcg_hmatrix(b200ks_ctx *c, double *H) { int k; (void)c; for (k = 0; k < 2 * s_eig_max * s_eig_max; k++) H[k] = 0.5 * k; return s_eig_max; }
int b200ks_eigcg_pairs(b200ks_ctx *c, double *val, int n) { int j; (void)c; logf_("eigcg_pairs %d\n", n); for (j = 0; j < n && j < s_eig_n; j++) val[j] = 0.001 * (j + 1); return s_eig_n; }
const char *b200ks_last_error(void) { return "stub error"; }
unsigned long long b200ks_fingerprint(const void *p, size_t bytes) {
  const unsigned char *b = (const unsigned char *)p;
  unsigned long long h = 1469598103934665603ull;
  size_t i;
  for (i = 0; i < bytes; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
int b200ks_load_links(b200ks_ctx *c, const void *fat, const void *lng, int host_prec, int long_recon) {
  (void)c; (void)fat; (void)lng;
  logf_("load_links prec %d recon %d\n", host_prec, long_recon);
  return 0;
}
/* the link cache of the real library without its concurrency: upload when told or when the pointers,
 * the precision or the content changed (every verifying mode compares right away) */
int b200ks_links_sync(b200ks_ctx *c, const void *fat, const void *lng, int host_prec, int changed_hint, int mode) {
  static const void *s_fat = NULL, *s_lng = NULL;
  static int s_prec = 0;
  static unsigned long long s_ff = 0, s_fl = 0;
  const size_t bytes = sites() * 72 * (host_prec == 2 ? 8 : 4);
  int load = changed_hint || fat != s_fat || lng != s_lng || host_prec != s_prec || mode == 1;
  unsigned long long ff = 0, fl = 0;
  if (mode >= 2 || load) {
    ff = b200ks_fingerprint(fat, bytes);
    fl = b200ks_fingerprint(lng, bytes);
    if (mode >= 2 && (ff != s_ff || fl != s_fl)) load = 1;
  }
  if (!load) return 0;
  s_fat = fat; s_lng = lng; s_prec = host_prec; s_ff = ff; s_fl = fl;
  return b200ks_load_links(c, fat, lng, host_prec, 0) < 0 ? -1 : 1;
}
static void fill(b200ks_invert_result *r, int iters) {
  memset(r, 0, sizeof(*r));
  r->final_rsq = 1e-20; r->final_relrsq = 0; r->size_r = 2e-20; r->size_relr = 0;
  r->final_iters = iters; r->final_restart = 1; r->converged = 1;
}
static void copy_parity(void *dst, const void *src, int parity, int host_prec) {
  const size_t vb = (size_t)6 * (host_prec == 2 ? 8 : 4), h = sites() / 2;
  if (parity & B200KS_EVEN) memcpy(dst, src, h * vb);
  if (parity & B200KS_ODD) memcpy((char *)dst + h * vb, (const char *)src + h * vb, h * vb);
}
int b200ks_congrad(b200ks_ctx *c, const void *src, void *dest, double mass, const b200ks_invert_args *a,
                   b200ks_invert_result *r, int host_prec) {
  (void)c;
  logf_("congrad mass %g parity %d max %d nrestart %d resid %g relresid %g mixed %d prec %d\n", mass, a->parity, a->max_iter,
        a->nrestart, a->resid, a->relresid, a->mixed_precision, host_prec);
  if (s_fail_next) { int e = s_fail_next; s_fail_next = 0; return e; }
  copy_parity(dest, src, a->parity, host_prec);
  fill(r, 17);
  return 17;
}
int b200ks_congrad_block(b200ks_ctx *c, int nsrc, const void *const *src, void *const *dest, double mass,
                         const b200ks_invert_args *a, b200ks_invert_result *r, int host_prec) {
  int k;
  (void)c;
  logf_("congrad_block nsrc %d mass %g parity %d\n", nsrc, mass, a->parity);
  for (k = 0; k < nsrc; k++) { copy_parity(dest[k], src[k], a->parity, host_prec); fill(&r[k], 10 + k); }
  return 10 * nsrc + nsrc * (nsrc - 1) / 2;
}
int b200ks_multicg(b200ks_ctx *c, const void *src, void *const *psim, const double *offsets, int n,
                   const b200ks_invert_args *a, b200ks_invert_result *r, int host_prec) {
  int j;
  (void)c;
  logf_("multicg n %d parity %d offsets", n, a->parity);
  for (j = 0; j < n; j++) { logf_(" %g", offsets[j]); copy_parity(psim[j], src, a->parity, host_prec); fill(&r[j], 23); }
  logf_("\n");
  return 23;
}
int b200ks_dslash(b200ks_ctx *c, const void *src, void *dest, int parity, int host_prec) {
  (void)c;
  logf_("dslash parity %d\n", parity);
  copy_parity(dest, src, parity, host_prec);
  return 0;
}
int b200ks_mat_invert_uml(b200ks_ctx *c, int nsrc, const void *const *src, void *const *dst, double mass,
                          const b200ks_invert_args *a, b200ks_invert_result *r, int host_prec) {
  int k;
  (void)c; (void)a;
  logf_("mat_invert_uml nsrc %d mass %g\n", nsrc, mass);
  for (k = 0; k < nsrc; k++) {
    copy_parity(dst[k], src[k], B200KS_EVENANDODD, host_prec);
    fill(&r[2 * k], 30); fill(&r[2 * k + 1], 3);
  }
  return 33 * nsrc;
}
int b200ks_vec_create(b200ks_ctx *c) { (void)c; s_live_vecs++; logf_("vec_create -> %d\n", s_next_vec); return s_next_vec++; }
int b200ks_vec_free(b200ks_ctx *c, int v) { (void)c; s_live_vecs--; logf_("vec_free %d\n", v); return 0; }
int b200ks_vec_upload(b200ks_ctx *c, int v, const void *host, int parity, int host_prec) {
  (void)c;
  logf_("vec_upload %d parity %d prec %d first %g\n", v, parity, host_prec,
        host_prec == 2 ? *(const double *)host : (double)*(const float *)host);
  return 0;
}
int b200ks_eig_set(b200ks_ctx *c, int n, const int *vecs, const double *eigval, int use_in_uml) {
  int j;
  (void)c;
  logf_("eig_set n %d uml %d", n, use_in_uml);
  for (j = 0; j < n; j++) logf_(" (%d %g)", vecs[j], eigval[j]);
  logf_("\n");
  return 0;
}
int b200ks_eig_use_in_uml(b200ks_ctx *c, int on) { (void)c; logf_("eig_use_in_uml %d\n", on); return 0; }

/* ---- the rest of what csrc/quda_shim.cu (route 2) calls ---------------------------------------------- */
int b200ks_device_count(void) { return 1; }
int b200ks_hisq_force(b200ks_ctx *c, int nterms, int num_naik_terms, const double *coeff, const void *const *multi_x,
                      const double *level2_coeff, const double *fat7_coeff, const void *wlink, const void *vlink,
                      const void *ulink, double eps, double force_filter, void *momentum, int host_prec) {
  int j;
  (void)c; (void)wlink; (void)vlink; (void)ulink; (void)momentum;
  logf_("hisq_force nterms %d naik %d eps %g filter %g prec %d coeff", nterms, num_naik_terms, eps, force_filter, host_prec);
  for (j = 0; j < 2 * (nterms + num_naik_terms); j++) logf_(" %g", coeff[j]);
  logf_(" x0");
  for (j = 0; j < nterms; j++) logf_(" %g", host_prec == 2 ? *(const double *)multi_x[j] : (double)*(const float *)multi_x[j]);
  logf_(" l2 %g %g f7 %g %g\n", level2_coeff[0], level2_coeff[5], fat7_coeff[0], fat7_coeff[2]);
  return 0;
}
int b200ks_ks_links(b200ks_ctx *c, const double *coeffs, const void *links, void *fat, void *lng, int host_prec) {
  (void)c; (void)links; (void)fat;
  logf_("ks_links c0 %g naik %g long %d prec %d\n", coeffs[0], coeffs[1], lng != NULL, host_prec);
  return 0;
}
int b200ks_unitarized_links(b200ks_ctx *c, const double *coeffs, const void *links, void *vlink, void *wlink, int host_prec,
                            long long *nsvd) {
  (void)c; (void)links; (void)wlink;
  logf_("unitarized_links c0 %g v %d prec %d\n", coeffs[0], vlink != NULL, host_prec);
  if (nsvd) *nsvd = 0;
  return 0;
}

/* meson tie-ups: corr[t][k] = (1000 spin + 10 t + k) + i (mx + 2 my + 4 mz + 0.125 (ex + ey + ez)) of momentum k */
int b200ks_meson_mom(b200ks_ctx *c, const void *antiquark, const void *quark, int host_prec, int spin, const int *r0, int nmom,
                     const int *mom, const char *mpar, double *corr) {
  int t, k;
  (void)c;
  if (s_fail_next) { int f = s_fail_next; s_fail_next = 0; return f; }
  logf_("meson_mom spin %d nmom %d r0 %d %d %d %d same %d prec %d\n", spin, nmom, r0[0], r0[1], r0[2], r0[3],
        antiquark == quark, host_prec);
  for (t = 0; t < s_dims[3]; t++)
    for (k = 0; k < nmom; k++) {
      corr[2 * (t * nmom + k)] = 1000.0 * spin + 10.0 * t + k;
      corr[2 * (t * nmom + k) + 1] = mom[3 * k] + 2 * mom[3 * k + 1] + 4 * mom[3 * k + 2] + 0.125 * (mpar[3 * k] + mpar[3 * k + 1] + mpar[3 * k + 2]);
    }
  return 0;
}
