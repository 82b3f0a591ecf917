/* oracle/ks_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C, double precision) of the reference's improved
 * staggered ("fat + Naik") Dirac hot path.  It is the parity checker for the
 * CUDA library: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it.  The product never links or calls it.
 *
 * PARITY PINNED: tests/test_oracle.py checks every function here against
 *   (1) oracle/_ref/libmilcref.so = the reference's own sources compiled from
 *       /root/reference (oracle/build_ref.sh), when present, and
 *   (2) tests/golden/ (npz files) = outputs of that same reference build, committed
 *       with the generator tests/golden/make_golden.py, so the check also runs
 *       where /root/reference is absent (the GPU box).
 *
 * What is restated (reference file:line):
 *   kso_node_index      generic/layout_hyper_prime.c:509-520
 *   kso_dslash          generic_ks/dslash_fn_dblstore.c:311-562 (maths),
 *                       generic_ks/dslash_fn.c:432-574 (adjoint-at-source form),
 *                       libraries/m_mv_s_4dir.c:139-254, m_amv_4vec.c:14-124
 *   kso_congrad         generic_ks/d_congrad5_fn_milc.c:60-407
 *   kso_multicg         generic_ks/ks_multicg_offset.c:63-505
 *   kso_relative_residue generic_ks/d_congrad5_fn_milc.c:37-56
 *   kso_deflate         generic_ks/mat_invert.c:131-183 (deflate + project_out), pinned through the
 *                       reference's deflated mat_invert_uml_field (tests/golden/make_golden_deflate.py)
 *   kso_meson_mom       generic_ks/ks_meson_mom.c:160-437 (site loops), generic_ks/spin_taste_ops.c:172-263
 *                       (local sink operators); pinned on the reference's ks_meson_cont_mom
 *                       (tests/golden/make_golden_meson.py)
 *
 * Data layout is MILC's host layout: site index i = node_index(x,y,z,t) (all even
 * sites, then all odd sites), vectors v[6*i + 2*c + {re,im}], links
 * L[(4*i + dir)*18 + (3*row + col)*2 + {re,im}], dir = X,Y,Z,T.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define KSO_EVEN 2
#define KSO_ODD 1
#define KSO_EVENANDODD 3

typedef struct {
  int n[4];
  long vol;
  /* neighbour tables in MILC site order: nb[k][i], k = 0..3 (+mu), 4..7 (-mu),
     8..11 (+3mu), 12..15 (-3mu) */
  int *nb[16];
} kso_geom;

static kso_geom G = {{0, 0, 0, 0}, 0, {0}};

/* generic/layout_hyper_prime.c:509-520 (single rank: squaresize == lattice) */
long kso_node_index(const int *n, int x, int y, int z, int t) {
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  long lex = x + (long)n[0] * (y + (long)n[1] * (z + (long)n[2] * t));
  if (((x + y + z + t) & 1) == 0) return lex / 2;
  return (lex + vol) / 2;
}

static void build_geom(const int *n) {
  int k, x[4], d;
  long i;
  if (G.vol && n[0] == G.n[0] && n[1] == G.n[1] && n[2] == G.n[2] && n[3] == G.n[3]) return;
  for (k = 0; k < 16; k++) { free(G.nb[k]); G.nb[k] = NULL; }
  memcpy(G.n, n, sizeof(G.n));
  G.vol = (long)n[0] * n[1] * n[2] * n[3];
  for (k = 0; k < 16; k++) G.nb[k] = (int *)malloc(sizeof(int) * G.vol);
  for (x[3] = 0; x[3] < n[3]; x[3]++)
    for (x[2] = 0; x[2] < n[2]; x[2]++)
      for (x[1] = 0; x[1] < n[1]; x[1]++)
        for (x[0] = 0; x[0] < n[0]; x[0]++) {
          i = kso_node_index(n, x[0], x[1], x[2], x[3]);
          for (d = 0; d < 4; d++) {
            /* periodic wrap: generic/com_vanilla.c:619-645, ks_spectrum/setup.c:1427-1445 */
            static const int hop[4] = {1, -1, 3, -3};
            int h;
            for (h = 0; h < 4; h++) {
              int y[4] = {x[0], x[1], x[2], x[3]};
              y[d] = (x[d] + hop[h] + 4 * n[d]) % n[d];
              G.nb[4 * h + d][i] = (int)kso_node_index(n, y[0], y[1], y[2], y[3]);
            }
          }
        }
}

/* c += A b          (libraries/m_mv_s_4dir.c:176-243: c_r = sum_k e[r][k] b_k) */
static void mv_acc(const double *A, const double *b, double *c, double sgn) {
  int r, k;
  for (r = 0; r < 3; r++) {
    double re = 0, im = 0;
    for (k = 0; k < 3; k++) {
      double ar = A[(3 * r + k) * 2], ai = A[(3 * r + k) * 2 + 1];
      re += ar * b[2 * k] - ai * b[2 * k + 1];
      im += ar * b[2 * k + 1] + ai * b[2 * k];
    }
    c[2 * r] += sgn * re;
    c[2 * r + 1] += sgn * im;
  }
}

/* c += sgn * A^dagger b   (libraries/m_amv_4vec.c:48-115: c_r = sum_k conj(e[k][r]) b_k) */
static void amv_acc(const double *A, const double *b, double *c, double sgn) {
  int r, k;
  for (r = 0; r < 3; r++) {
    double re = 0, im = 0;
    for (k = 0; k < 3; k++) {
      double ar = A[(3 * k + r) * 2], ai = -A[(3 * k + r) * 2 + 1];
      re += ar * b[2 * k] - ai * b[2 * k + 1];
      im += ar * b[2 * k + 1] + ai * b[2 * k];
    }
    c[2 * r] += sgn * re;
    c[2 * r + 1] += sgn * im;
  }
}

static void parity_range(long vol, int parity, long *lo, long *hi) {
  long vh = vol / 2;
  *lo = (parity == KSO_ODD) ? vh : 0;
  *hi = (parity == KSO_EVEN) ? vh : vol;
}

/* dest(x) = sum_mu [ F_mu(x) src(x+mu) + L_mu(x) src(x+3mu)
 *                  - F_mu(x-mu)^+ src(x-mu) - L_mu(x-3mu)^+ src(x-3mu) ]
 * for x of `parity`; other-parity entries of dest are untouched and src == dest
 * is legal (generic_ks/d_congrad5_fn_milc.c:197).  Backward hops use the adjoint
 * of the link stored AT THE NEIGHBOUR (generic_ks/fn_links_milc.c:132-146,180-194).
 */
void kso_dslash(const int *n, const double *fat, const double *lng, const double *src,
                double *dest, int parity) {
  long lo, hi, i;
  build_geom(n);
  parity_range(G.vol, parity, &lo, &hi);
  /* src==dest is legal only because parities differ; with EVENANDODD they
     would alias, so stage the result. */
  double *out = dest;
  if (parity == KSO_EVENANDODD && src == dest) out = (double *)malloc(sizeof(double) * 6 * G.vol);
#pragma omp parallel for
  for (i = lo; i < hi; i++) {
    double acc[6] = {0, 0, 0, 0, 0, 0};
    int d;
    for (d = 0; d < 4; d++) mv_acc(fat + (4 * i + d) * 18, src + 6 * (long)G.nb[d][i], acc, 1.0);
    for (d = 0; d < 4; d++) mv_acc(lng + (4 * i + d) * 18, src + 6 * (long)G.nb[8 + d][i], acc, 1.0);
    for (d = 0; d < 4; d++) {
      long j = G.nb[4 + d][i];
      amv_acc(fat + (4 * j + d) * 18, src + 6 * j, acc, -1.0);
    }
    for (d = 0; d < 4; d++) {
      long j = G.nb[12 + d][i];
      amv_acc(lng + (4 * j + d) * 18, src + 6 * j, acc, -1.0);
    }
    memcpy(out + 6 * i, acc, sizeof(acc));
  }
  if (out != dest) {
    memcpy(dest, out, sizeof(double) * 6 * G.vol);
    free(out);
  }
}

static double dot_re(const double *a, const double *b, long lo, long hi) {
  double s = 0;
  long i;
#pragma omp parallel for reduction(+ : s)
  for (i = 6 * lo; i < 6 * hi; i++) s += a[i] * b[i];
  return s;
}

/* generic_ks/d_congrad5_fn_milc.c:37-56 */
double kso_relative_residue(const int *n, const double *p, const double *q, int parity) {
  long lo, hi, i;
  double residue = 0;
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  parity_range(vol, parity, &lo, &hi);
  for (i = lo; i < hi; i++) {
    double num = 0, den = 0;
    int k;
    for (k = 0; k < 6; k++) { num += p[6 * i + k] * p[6 * i + k]; den += q[6 * i + k] * q[6 * i + k]; }
    residue += (den == 0) ? 1.0 : (num / den);
  }
  if (parity == KSO_EVENANDODD) return sqrt(residue / vol);
  return sqrt(2 * residue / vol);
}

/* out[0..6] = final_rsq, final_relrsq, size_r, size_relr, final_iters, final_restart, converged */

/* Single-mass CG on one parity, generic_ks/d_congrad5_fn_milc.c:60-407.
 * Solves (4 m^2 - D_pp' D_p'p) dest = src; dest holds the initial guess.
 * fewsums != 0 follows the reference's default -DFEWSUMS arithmetic (:282-308,339):
 * the recursive |r|^2 is oldrsq + 2a<ttt|r> + a^2<ttt|ttt> instead of a direct sum.
 */
int kso_congrad(const int *n, const double *fat, const double *lng, const double *src,
                double *dest, double mass, int parity, int niter, int max_restarts,
                double resid, double relresid, int fewsums, double *out) {
  long lo, hi, i, vol;
  int otherparity = (parity == KSO_EVEN) ? KSO_ODD : KSO_EVEN;
  double rsqmin = resid * resid, relrsqmin = relresid * relresid;
  double msq_x4 = 4.0 * mass * mass;
  double source_norm, rsq = 0, relrsq = 1.0, oldrsq, pkp, a, b;
  double actual_rsq = 999., c_tr = 0, c_tt = 0;
  double size_r = 0, size_relr = 1.0, final_rsq = 0, final_relrsq = 0;
  int iteration = 0, nrestart = 0, max_cg = max_restarts * niter, converged = 1;
  double *ttt, *cg_p, *res;
  build_geom(n);
  vol = G.vol;
  parity_range(vol, parity, &lo, &hi);
  memset(out, 0, 7 * sizeof(double));
  out[6] = 1;

  source_norm = dot_re(src, src, lo, hi);
  if (source_norm == 0.0) { /* :136-152 trivial solution */
    memset(dest + 6 * lo, 0, sizeof(double) * 6 * (hi - lo));
    out[3] = 1.0;
    return 0;
  }
  ttt = (double *)calloc(6 * vol, sizeof(double));
  cg_p = (double *)calloc(6 * vol, sizeof(double));
  res = (double *)calloc(6 * vol, sizeof(double));

  for (;;) {
    if ((iteration % niter == 0) ||
        ((rsqmin <= 0 || rsqmin > size_r) && (relrsqmin <= 0 || relrsqmin > size_relr))) {
      /* (re)start from the true residual, :177-240 */
      kso_dslash(n, fat, lng, dest, ttt, otherparity);
      kso_dslash(n, fat, lng, ttt, ttt, parity);
      rsq = 0;
      for (i = 6 * lo; i < 6 * hi; i++) {
        ttt[i] = ttt[i] - msq_x4 * dest[i];
        res[i] = src[i] + ttt[i];
        cg_p[i] = res[i];
        rsq += res[i] * res[i];
      }
      actual_rsq = rsq;
      if (relrsqmin > 0) relrsq = kso_relative_residue(n, res, dest, parity);
      final_rsq = rsq / source_norm;
      final_relrsq = relrsq;
      iteration++;
      if (iteration >= max_cg || nrestart >= max_restarts ||
          ((rsqmin <= 0 || rsqmin > final_rsq) && (relrsqmin <= 0 || relrsqmin > final_relrsq)))
        break;
      nrestart++;
    }
    oldrsq = fewsums ? actual_rsq : rsq;
    kso_dslash(n, fat, lng, cg_p, ttt, otherparity);
    kso_dslash(n, fat, lng, ttt, ttt, parity);
    pkp = 0; c_tr = 0; c_tt = 0;
    for (i = 6 * lo; i < 6 * hi; i++) {
      ttt[i] = ttt[i] - msq_x4 * cg_p[i];
      pkp += cg_p[i] * ttt[i];
      c_tr += ttt[i] * res[i];
      c_tt += ttt[i] * ttt[i];
    }
    iteration++;
    a = -rsq / pkp;
    actual_rsq = 0;
    for (i = 6 * lo; i < 6 * hi; i++) {
      dest[i] += a * cg_p[i];
      res[i] += a * ttt[i];
      actual_rsq += res[i] * res[i];
    }
    rsq = fewsums ? (oldrsq + 2.0 * a * c_tr + a * a * c_tt) : actual_rsq;
    if (relrsqmin > 0) relrsq = kso_relative_residue(n, res, dest, parity);
    size_r = rsq / source_norm;
    size_relr = relrsq;
    b = rsq / oldrsq;
    for (i = 6 * lo; i < 6 * hi; i++) cg_p[i] = res[i] + b * cg_p[i];
  }
  if (nrestart == max_restarts || iteration == max_cg) converged = 0;
  out[0] = final_rsq; out[1] = final_relrsq; out[2] = size_r; out[3] = size_relr;
  out[4] = iteration; out[5] = nrestart; out[6] = converged;
  free(ttt); free(cg_p); free(res);
  return iteration;
}

/* Multi-shift CG, generic_ks/ks_multicg_offset.c:63-505 (one parity per call;
 * the reference's EVENANDODD is "EVEN then ODD" and is done by the caller).
 * psim: num_offsets fields of 6*vol doubles, zeroed here (initial guess ignored, :230).
 * out: 7 doubles per offset, as kso_congrad.
 */
int kso_multicg(const int *n, const double *fat, const double *lng, const double *src,
                double *psim, const double *offsets, int num_offsets, int parity, int max,
                int nrest, double resid, double relresid, double *out) {
  long lo, hi, i, vol;
  int otherparity = (parity == KSO_EVEN) ? KSO_ODD : KSO_EVEN;
  int niter = max * nrest, iteration = 0, j, j_low = -1, num_offsets_now = num_offsets;
  double rsqmin = resid * resid, relrsqmin = relresid * relresid;
  double source_norm, rsq, oldrsq, pkp, rsqstop, relrsq = 0, c1, c2, offset_low = 1.0e+20, shift0;
  double *shifts, *zeta_i, *zeta_im1, *zeta_ip1, *beta_i, *beta_im1, *alpha, *ttt, *cg_p, *res;
  double **pm;
  int converged = 0;
  if (num_offsets == 0) return 0;
  build_geom(n);
  vol = G.vol;
  parity_range(vol, parity, &lo, &hi);
  memset(out, 0, 7 * num_offsets * sizeof(double));
  for (j = 0; j < num_offsets; j++) out[7 * j + 6] = 1;

  shifts = (double *)malloc(num_offsets * sizeof(double));
  zeta_i = (double *)malloc(num_offsets * sizeof(double));
  zeta_im1 = (double *)malloc(num_offsets * sizeof(double));
  zeta_ip1 = (double *)malloc(num_offsets * sizeof(double));
  beta_i = (double *)malloc(num_offsets * sizeof(double));
  beta_im1 = (double *)malloc(num_offsets * sizeof(double));
  alpha = (double *)malloc(num_offsets * sizeof(double));
  pm = (double **)malloc(num_offsets * sizeof(double *));
  for (j = 0; j < num_offsets; j++) { /* :181-194 */
    shifts[j] = offsets[j];
    if (offsets[j] < offset_low) { offset_low = offsets[j]; j_low = j; }
  }
  for (j = 0; j < num_offsets; j++) {
    pm[j] = (double *)calloc(6 * vol, sizeof(double));
    if (j != j_low) shifts[j] -= shifts[j_low];
  }
  shift0 = -shifts[j_low];
  ttt = (double *)calloc(6 * vol, sizeof(double));
  cg_p = (double *)calloc(6 * vol, sizeof(double));
  res = (double *)calloc(6 * vol, sizeof(double));

  source_norm = dot_re(src, src, lo, hi);
  for (i = 6 * lo; i < 6 * hi; i++) {
    res[i] = src[i];
    cg_p[i] = src[i];
    for (j = 0; j < num_offsets; j++) { psim[(long)j * 6 * vol + i] = 0; pm[j][i] = src[i]; }
  }
  rsq = source_norm;
  if (source_norm == 0.0) goto done_trivial;

  iteration++;
  rsqstop = rsqmin * source_norm;
  for (j = 0; j < num_offsets; j++) { zeta_im1[j] = zeta_i[j] = 1.0; beta_im1[j] = -1.0; alpha[j] = 0.0; }

  do {
    oldrsq = rsq;
    kso_dslash(n, fat, lng, cg_p, ttt, otherparity);
    kso_dslash(n, fat, lng, ttt, ttt, parity);
    pkp = 0;
    for (i = 6 * lo; i < 6 * hi; i++) { ttt[i] += shift0 * cg_p[i]; pkp += cg_p[i] * ttt[i]; }
    iteration++;
    beta_i[j_low] = -rsq / pkp;
    zeta_ip1[j_low] = 1.0;
    for (j = 0; j < num_offsets_now; j++) if (j != j_low) { /* :327-355 */
      zeta_ip1[j] = zeta_i[j] * zeta_im1[j] * beta_im1[j_low];
      c1 = beta_i[j_low] * alpha[j_low] * (zeta_im1[j] - zeta_i[j]);
      c2 = zeta_im1[j] * beta_im1[j_low] * (1.0 + shifts[j] * beta_i[j_low]);
      if (c1 + c2 != 0.0) zeta_ip1[j] /= c1 + c2; else zeta_ip1[j] = 0.0;
      if (zeta_i[j] != 0.0) beta_i[j] = beta_i[j_low] * zeta_ip1[j] / zeta_i[j];
      else {
        zeta_ip1[j] = 0.0; beta_i[j] = 0.0;
        if (j == num_offsets_now - 1 && j > j_low) num_offsets_now--;
      }
    }
    rsq = 0;
    for (i = 6 * lo; i < 6 * hi; i++) {
      for (j = 0; j < num_offsets_now; j++) psim[(long)j * 6 * vol + i] += beta_i[j] * pm[j][i];
      res[i] += beta_i[j_low] * ttt[i];
      rsq += res[i] * res[i];
    }
    if (relrsqmin > 0) {
      relrsq = 0;
      for (j = 0; j < num_offsets_now; j++) {
        double r = kso_relative_residue(n, res, psim + (long)j * 6 * vol, parity);
        out[7 * j + 1] = r;
        if (r > relrsq) relrsq = r;
      }
    }
    if ((rsqstop > 0 && rsq <= rsqstop) || (relrsqmin > 0 && relrsq <= relrsqmin)) { converged = 1; break; }
    alpha[j_low] = rsq / oldrsq;
    for (j = 0; j < num_offsets_now; j++) if (j != j_low) { /* :431-444 */
      if (zeta_i[j] * beta_i[j_low] != 0.0)
        alpha[j] = alpha[j_low] * zeta_ip1[j] * beta_i[j] / (zeta_i[j] * beta_i[j_low]);
      else alpha[j] = 0.0;
    }
    for (i = 6 * lo; i < 6 * hi; i++) {
      for (j = 0; j < num_offsets_now; j++) pm[j][i] = zeta_ip1[j] * res[i] + alpha[j] * pm[j][i];
      cg_p[i] = pm[j_low][i];
    }
    for (j = 0; j < num_offsets_now; j++) { beta_im1[j] = beta_i[j]; zeta_im1[j] = zeta_i[j]; zeta_i[j] = zeta_ip1[j]; }
  } while (iteration < niter);

  for (j = 0; j < num_offsets; j++) {
    out[7 * j + 0] = rsq / source_norm;
    out[7 * j + 2] = out[7 * j + 0];
    out[7 * j + 3] = out[7 * j + 1];
    out[7 * j + 4] = iteration;
    out[7 * j + 6] = converged;
  }
done_trivial:
  if (source_norm == 0.0)
    for (j = 0; j < num_offsets; j++) out[7 * j + 4] = iteration;
  for (j = 0; j < num_offsets; j++) free(pm[j]);
  free(pm); free(shifts); free(zeta_i); free(zeta_im1); free(zeta_ip1);
  free(beta_i); free(beta_im1); free(alpha); free(ttt); free(cg_p); free(res);
  return iteration;
}

/* ---- low-mode deflation of a trial solution -------------------------
 * deflate() + project_out(), generic_ks/mat_invert.c:131-183: on the sites of `parity`
 *     dst <- dst - sum_j v_j <v_j|dst>            (one vector after the other, j = nvecs-1 .. 0)
 *     dst <- dst + sum_j v_j <v_j|src> / (eigval_j + 4 m^2)
 * eigvec: nvecs fields of V colour vectors (both parities filled, orthonormal on each parity),
 * eigval: the eigenvalues of -D_eo D_oe they belong to.  Gives mat_invert_uml_field / _cg_field
 * (:186-257,328-402) the exact solution in the span of the vectors as the CG's starting point. */
void kso_deflate(const int *n, double *dst, const double *src, double mass, int nvecs, const double *eigvec,
                 const double *eigval, int parity) {
  const long vol = (long)n[0] * n[1] * n[2] * n[3];
  long lo, hi, i;
  int j, c;
  parity_range(vol, parity, &lo, &hi);
  for (j = nvecs - 1; j >= 0; j--) {
    const double *v = eigvec + (size_t)j * vol * 6;
    double re = 0, im = 0;
    for (i = lo; i < hi; i++)
      for (c = 0; c < 3; c++) { /* conj(v) * dst */
        const double vr = v[6 * i + 2 * c], vi = v[6 * i + 2 * c + 1], dr = dst[6 * i + 2 * c], di = dst[6 * i + 2 * c + 1];
        re += vr * dr + vi * di;
        im += vr * di - vi * dr;
      }
    for (i = lo; i < hi; i++)
      for (c = 0; c < 3; c++) {
        const double vr = v[6 * i + 2 * c], vi = v[6 * i + 2 * c + 1];
        dst[6 * i + 2 * c] -= re * vr - im * vi;
        dst[6 * i + 2 * c + 1] -= re * vi + im * vr;
      }
  }
  for (j = 0; j < nvecs; j++) {
    const double *v = eigvec + (size_t)j * vol * 6;
    const double den = eigval[j] + 4.0 * mass * mass;
    double re = 0, im = 0;
    for (i = lo; i < hi; i++)
      for (c = 0; c < 3; c++) {
        const double vr = v[6 * i + 2 * c], vi = v[6 * i + 2 * c + 1], sr = src[6 * i + 2 * c], si = src[6 * i + 2 * c + 1];
        re += vr * sr + vi * si;
        im += vr * si - vi * sr;
      }
    re /= den; im /= den;
    for (i = lo; i < hi; i++)
      for (c = 0; c < 3; c++) {
        const double vr = v[6 * i + 2 * c], vi = v[6 * i + 2 * c + 1];
        dst[6 * i + 2 * c] += re * vr - im * vi;
        dst[6 * i + 2 * c + 1] += re * vi + im * vr;
      }
  }
}

/* ---- meson tie-ups ---------------------------------------------------
 * The site loops of ks_meson_cont_mom (generic_ks/ks_meson_mom.c:160-437) for ONE sink spin-taste
 * assignment with a LOCAL sink operator (spin >= 0: local(), generic_ks/spin_taste_ops.c:172-263,
 * applied to the antiquark) or none (spin < 0: the caller has applied it):
 *   ftfact_p(x) = ff(2 pi/nx (x - r0x) px, ex, ff-chain ...)          (:137-157, :262-283)
 *   meson(x)    = su3_dot(antiquark(x), quark(x))                     (:349-363)
 *   corr[t][p] += meson(x) ftfact_p(x)  over the sites of slice t, in node-index order (:364-383)
 * mom[3p + d], mpar[3p + d] (KSO_EVEN / KSO_ODD / KSO_EVENANDODD) as q_momstore / q_parity;
 * corr[(t*nmom + p)*2 + {re,im}] is overwritten.  norm_v and the prop[] accumulation stay with the caller. */
static void kso_ff(double theta, int parity, double *re, double *im) {
  const double tr = *re, ti = *im;
  if (parity == KSO_EVEN) { *re = tr * cos(theta); *im = ti * cos(theta); }
  else if (parity == KSO_ODD) { *re = -ti * sin(theta); *im = tr * sin(theta); }
  else { *re = tr * cos(theta) - ti * sin(theta); *im = ti * cos(theta) + tr * sin(theta); }
}

void kso_meson_mom(const int *n, const double *antiquark, const double *quark, int spin, const int *r0, int nmom,
                   const int *mom, const char *mpar, double *corr) {
  const double PI_ = 3.14159265358979323846;
  const double fx = 2.0 * PI_ / (1.0 * n[0]), fy = 2.0 * PI_ / (1.0 * n[1]), fz = 2.0 * PI_ / (1.0 * n[2]);
  int x, y, z, t, p, c;
  memset(corr, 0, sizeof(double) * 2 * (size_t)n[3] * nmom);
  for (t = 0; t < n[3]; t++) for (z = 0; z < n[2]; z++) for (y = 0; y < n[1]; y++) for (x = 0; x < n[0]; x++) {
    const long i = kso_node_index(n, x, y, z, t);
    double re = 0, im = 0, sign = 1.0;
    for (c = 0; c < 3; c++) { /* su3_dot: conj(a) * b */
      const double ar = antiquark[6 * i + 2 * c], ai = antiquark[6 * i + 2 * c + 1];
      const double br = quark[6 * i + 2 * c], bi = quark[6 * i + 2 * c + 1];
      re += ar * br + ai * bi;
      im += ar * bi - ai * br;
    }
    if (spin >= 0) {
      /* spin_sign(): for each gamma_mu bit a factor (-)^(x_mu - r0_mu) eps(x - r0); antiquark_sign_flip(): eps */
      const int h[4] = {(x - r0[0]) & 1, (y - r0[1]) & 1, (z - r0[2]) & 1, (t - r0[3]) & 1};
      const int hp = (h[0] + h[1] + h[2] + h[3]) & 1;
      int j;
      for (j = 0; j < 4; j++) if ((spin & (1 << j)) && (hp ^ h[j])) sign = -sign;
      if (hp) sign = -sign;
    }
    re *= sign; im *= sign;
    for (p = 0; p < nmom; p++) {
      double fr = 1.0, fi = 0.0;
      kso_ff(fx * (x - r0[0]) * mom[3 * p + 0], mpar[3 * p + 0], &fr, &fi);
      kso_ff(fy * (y - r0[1]) * mom[3 * p + 1], mpar[3 * p + 1], &fr, &fi);
      kso_ff(fz * (z - r0[2]) * mom[3 * p + 2], mpar[3 * p + 2], &fr, &fi);
      corr[(t * (size_t)nmom + p) * 2] += re * fr - im * fi;
      corr[(t * (size_t)nmom + p) * 2 + 1] += re * fi + im * fr;
    }
  }
}
