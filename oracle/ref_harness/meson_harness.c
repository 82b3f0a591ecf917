/* oracle/ref_harness/meson_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * ctypes-friendly face of the reference's meson tie-ups (generic_ks/ks_meson_mom.c, spin_taste_ops.c and
 * generic_wilson/gammas.c compiled unmodified from /root/reference by oracle/build_ref.sh together with harness.c).
 *   milcref_spin_taste_index   spin_taste_index()    (spin_taste_ops.c:1183)
 *   milcref_spin_taste_op      spin_taste_op_fn()    (sink operator applied to a field, fat/long links of harness.c)
 *   milcref_meson_cont_mom     ks_meson_cont_mom()   (ks_meson_mom.c:160-437) for ncorr correlators; the grouping by
 *                              sink spin-taste assignment (corr_table / num_corr_mom) is built here the way
 *                              ks_spectrum's input parser does: consecutive correlators with the same operator
 */
#include "generic_ks_includes.h"
#include <string.h>

extern fn_links_t *milcref_fn(void);

/* ape_links (ks_spectrum's smeared links, ks_spectrum/lattice.h): spin_taste_op() rephases them around EVERY operator,
 * local ones included, and the one-link operators shift with them.  Any field of 3 x 3 matrices will do for a test. */
int milcref_set_ape_links(const Real *links) {
  if (ape_links == NULL) ape_links = (su3_matrix *)malloc(4 * sites_on_node * sizeof(su3_matrix));
  memcpy(ape_links, links, 4 * sites_on_node * sizeof(su3_matrix));
  return 0;
}

int milcref_spin_taste_index(const char *label) { return spin_taste_index((char *)label); }

void milcref_spin_taste_op(int index, const int *r0, Real *dest, const Real *src) {
  int r[4] = {r0[0], r0[1], r0[2], r0[3]};
  spin_taste_op_fn(milcref_fn(), index, r, (su3_vector *)dest, (su3_vector *)src);
}

/* prop[m][t] (re, im) is ACCUMULATED like the reference does */
int milcref_meson_cont_mom(const Real *src1, const Real *src2, int nmom, const int *mom, const char *mpar, int ncorr,
                           const int *spin_taste, const int *p_index, const int *phase, const double *factor,
                           const int *corr_index, int nprop, const int *r0, double *prop) {
  int p, c, g, m, t, ng = 0;
  int **q_momstore = (int **)malloc(nmom * sizeof(int *));
  char **q_parity = (char **)malloc(nmom * sizeof(char *));
  int *num_corr_mom = (int *)calloc(ncorr, sizeof(int));
  int **corr_table = (int **)malloc(ncorr * sizeof(int *));
  Real *fac = (Real *)malloc(ncorr * sizeof(Real));
  complex **pr = (complex **)malloc(nprop * sizeof(complex *));
  int r[4] = {r0[0], r0[1], r0[2], r0[3]};
  for (p = 0; p < nmom; p++) {
    q_momstore[p] = (int *)malloc(3 * sizeof(int));
    q_parity[p] = (char *)malloc(3);
    for (c = 0; c < 3; c++) { q_momstore[p][c] = mom[3 * p + c]; q_parity[p][c] = mpar[3 * p + c]; }
  }
  for (c = 0; c < ncorr; c++) {
    fac[c] = (Real)factor[c];
    if (c == 0 || spin_taste[c] != spin_taste[c - 1]) {
      corr_table[ng] = (int *)malloc(ncorr * sizeof(int));
      num_corr_mom[ng] = 0;
      ng++;
    }
    g = ng - 1;
    corr_table[g][num_corr_mom[g]++] = c;
  }
  for (m = 0; m < nprop; m++) {
    pr[m] = (complex *)malloc(nt * sizeof(complex));
    for (t = 0; t < nt; t++) { pr[m][t].real = prop[(m * nt + t) * 2]; pr[m][t].imag = prop[(m * nt + t) * 2 + 1]; }
  }
  ks_meson_cont_mom(pr, (su3_vector *)src1, (su3_vector *)src2, nmom, q_momstore, q_parity, ng, num_corr_mom, corr_table,
                    (int *)p_index, milcref_fn(), milcref_fn(), (int *)spin_taste, (int *)phase, fac, (int *)corr_index, r);
  for (m = 0; m < nprop; m++) {
    for (t = 0; t < nt; t++) { prop[(m * nt + t) * 2] = pr[m][t].real; prop[(m * nt + t) * 2 + 1] = pr[m][t].imag; }
    free(pr[m]);
  }
  for (p = 0; p < nmom; p++) { free(q_momstore[p]); free(q_parity[p]); }
  for (g = 0; g < ng; g++) free(corr_table[g]);
  free(q_momstore); free(q_parity); free(num_corr_mom); free(corr_table); free(fac); free(pr);
  return ng;
}
