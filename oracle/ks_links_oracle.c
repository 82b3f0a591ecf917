/* oracle/ks_links_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C, double precision) of the reference's HISQ/asqtad fermion-link
 * construction.  Parity checker for the CUDA link kernels: only
 * tests/ and bench.py's cpu_baseline leg may load it.
 *
 * PARITY PINNED: tests/test_oracle.py checks every function here against the reference's own
 * compiled chain (oracle/_ref/libmilcref.so: create_hisq_links_milc, load_fatlinks_cpu,
 * load_lnglinks, u3_unitarize_analytic through oracle/ref_harness/harness.c) and against the
 * committed golden tests/golden/ref_hisq_links.npz that build produced.
 *
 * What is restated (reference file:line):
 *   ksl_smear       generic_ks/fermion_links_fn_load_milc.c:120-275 (load_fatlinks_cpu,
 *                   ASQ_OPTIMIZED_FATTENING branch) with generic/general_staple.c:41-123
 *                   (compute_gen_staple_field), and :45-107 (load_lnglinks) for the straight
 *                   three-link Naik path
 *   ksl_unitarize   generic_ks/su3_mat_op.c:828-1205 (u3_unitarize_analytic, double branch);
 *                   its SVD branch (svd3x3, taken when the Cayley-Hamilton eigenvalues fail the
 *                   determinant check) is restated as a one-sided Jacobi SVD: any SVD V = A S B^+
 *                   gives the same unitary factor A B^+
 *   ksl_hisq_links  generic_ks/fermion_links_hisq_load_milc.c:531-586 (U -> V -> Y = W -> X),
 *                   single Naik epsilon = 0
 *
 * Layout: MILC host layout, site index i = node_index(x,y,z,t) (even sites, then odd sites),
 * links L[(4*i + dir)*18 + (3*row + col)*2 + {re,im}].  KS phases and boundary signs are part
 * of the input links (phases_in = 1).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double e[3][3][2]; } mat;

static long node_index(const int *n, int x, int y, int z, int t) {
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  long lex = x + (long)n[0] * (y + (long)n[1] * (z + (long)n[2] * t));
  return (((x + y + z + t) & 1) == 0) ? lex / 2 : (lex + vol) / 2;
}

/* nb[d][i] = site at +1 in direction d (d < 4) or -1 in direction d-4 */
static int *build_nb(const int *n, int d) {
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  int *nb = (int *)malloc(sizeof(int) * vol);
  int x[4];
  for (x[3] = 0; x[3] < n[3]; x[3]++)
    for (x[2] = 0; x[2] < n[2]; x[2]++)
      for (x[1] = 0; x[1] < n[1]; x[1]++)
        for (x[0] = 0; x[0] < n[0]; x[0]++) {
          int y[4] = {x[0], x[1], x[2], x[3]};
          int mu = d & 3;
          y[mu] = (x[mu] + (d < 4 ? 1 : -1) + n[mu]) % n[mu];
          nb[node_index(n, x[0], x[1], x[2], x[3])] = (int)node_index(n, y[0], y[1], y[2], y[3]);
        }
  return nb;
}

/* c = a b   (libraries/m_mat_nn.c) */
static void mult_nn(const mat *a, const mat *b, mat *c) {
  int i, j, k;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (k = 0; k < 3; k++) {
        re += a->e[i][k][0] * b->e[k][j][0] - a->e[i][k][1] * b->e[k][j][1];
        im += a->e[i][k][0] * b->e[k][j][1] + a->e[i][k][1] * b->e[k][j][0];
      }
      c->e[i][j][0] = re;
      c->e[i][j][1] = im;
    }
}
/* c = a b^dagger   (libraries/m_mat_na.c) */
static void mult_na(const mat *a, const mat *b, mat *c) {
  int i, j, k;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (k = 0; k < 3; k++) {
        re += a->e[i][k][0] * b->e[j][k][0] + a->e[i][k][1] * b->e[j][k][1];
        im += a->e[i][k][1] * b->e[j][k][0] - a->e[i][k][0] * b->e[j][k][1];
      }
      c->e[i][j][0] = re;
      c->e[i][j][1] = im;
    }
}
/* c = a^dagger b   (libraries/m_mat_an.c) */
static void mult_an(const mat *a, const mat *b, mat *c) {
  int i, j, k;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (k = 0; k < 3; k++) {
        re += a->e[k][i][0] * b->e[k][j][0] + a->e[k][i][1] * b->e[k][j][1];
        im += a->e[k][i][0] * b->e[k][j][1] - a->e[k][i][1] * b->e[k][j][0];
      }
      c->e[i][j][0] = re;
      c->e[i][j][1] = im;
    }
}
/* a += s b */
static void axpy(mat *a, double s, const mat *b) {
  int i, j, r;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++)
      for (r = 0; r < 2; r++) a->e[i][j][r] += s * b->e[i][j][r];
}

/* generic/general_staple.c:41-123.  `link` holds the mu link as `stride` matrices per site. */
static void gen_staple(long vol, int *const *nb, mat *staple, int mu, int nu, const mat *link, int stride, mat *fat,
                       double coef, const mat *links) {
  long i;
  mat *tempmat = (mat *)malloc(sizeof(mat) * vol);
  for (i = 0; i < vol; i++) { /* upper staple: U_nu(x) link(x+nu) U_nu(x+mu)^+ */
    mat t1, t2;
    mult_na(&link[(long)stride * nb[nu][i]], &links[4l * nb[mu][i] + nu], &t1);
    mult_nn(&links[4 * i + nu], &t1, &t2);
    if (staple) staple[i] = t2;
    else axpy(&fat[4 * i + mu], coef, &t2);
  }
  for (i = 0; i < vol; i++) { /* lower staple, built at x-nu: U_nu^+ link U_nu(.+mu) */
    mat t1;
    mult_an(&links[4 * i + nu], &link[(long)stride * i], &t1);
    mult_nn(&t1, &links[4l * nb[mu][i] + nu], &tempmat[i]);
  }
  for (i = 0; i < vol; i++) {
    const mat *low = &tempmat[nb[4 + nu][i]];
    if (staple) {
      axpy(&staple[i], 1.0, low);
      axpy(&fat[4 * i + mu], coef, &staple[i]);
    } else {
      axpy(&fat[4 * i + mu], coef, low);
    }
  }
  free(tempmat);
}

/* coeffs = {one_link, naik, three_staple, five_staple, seven_staple, lepage}; lng may be NULL */
void ksl_smear(const int *n, const double *coeffs, const double *links_, double *fat_, double *lng_) {
  const long vol = (long)n[0] * n[1] * n[2] * n[3];
  const mat *links = (const mat *)links_;
  mat *fat = (mat *)fat_, *lng = (mat *)lng_;
  const double one_link = coeffs[0], naik = coeffs[1], three = coeffs[2], five = coeffs[3], seven = coeffs[4],
               lepage = coeffs[5];
  int *nb[8];
  int d, dir, nu, rho, sig;
  long i;
  mat *staple = (mat *)malloc(sizeof(mat) * vol), *tempmat1 = (mat *)malloc(sizeof(mat) * vol);
  for (d = 0; d < 8; d++) nb[d] = build_nb(n, d);
  /* fermion_links_fn_load_milc.c:214-256 */
  for (dir = 0; dir < 4; dir++) {
    const double c1 = one_link - 6.0 * lepage;
    for (i = 0; i < vol; i++) {
      memset(&fat[4 * i + dir], 0, sizeof(mat));
      axpy(&fat[4 * i + dir], c1, &links[4 * i + dir]);
    }
    if (three == 0.0 && lepage == 0.0 && five == 0.0) continue;
    for (nu = 0; nu < 4; nu++) {
      if (nu == dir) continue;
      gen_staple(vol, nb, staple, dir, nu, links + dir, 4, fat, three, links);
      gen_staple(vol, nb, NULL, dir, nu, staple, 1, fat, lepage, links);
      for (rho = 0; rho < 4; rho++) {
        if (rho == dir || rho == nu) continue;
        gen_staple(vol, nb, tempmat1, dir, rho, staple, 1, fat, five, links);
        for (sig = 0; sig < 4; sig++) {
          if (sig == dir || sig == nu || sig == rho) continue;
          gen_staple(vol, nb, NULL, dir, sig, tempmat1, 1, fat, seven, links);
        }
      }
    }
  }
  /* fermion_links_fn_load_milc.c:45-107 for the straight three-link path: the backward path
     product is adjointed and weighted with -coeff, i.e. lng = naik * U(x) U(x+mu) U(x+2mu) */
  if (lng)
    for (dir = 0; dir < 4; dir++)
      for (i = 0; i < vol; i++) {
        const long i1 = nb[dir][i], i2 = nb[dir][i1];
        mat t1, t2;
        mult_nn(&links[4 * i + dir], &links[4 * i1 + dir], &t1);
        mult_nn(&t1, &links[4 * i2 + dir], &t2);
        memset(&lng[4 * i + dir], 0, sizeof(mat));
        axpy(&lng[4 * i + dir], naik, &t2);
      }
  for (d = 0; d < 8; d++) free(nb[d]);
  free(staple);
  free(tempmat1);
}

/* ---- U(3) projection ------------------------------------------------------------------------ */
#define KSL_EPS 1.0e-14 /* U3_UNIT_ANALYTIC_EPS, include/su3_mat_op.h:18 */

/* One-sided (Hestenes) Jacobi SVD of a complex 3x3 matrix: rotations from the right make the
   columns orthogonal, A J = U S; the unitary polar factor is U J^+ .  Stands in for svd3x3. */
static void polar_by_svd(const mat *V, mat *W) {
  double a[3][3][2], v[3][3][2];
  int i, j, p, q, sweep;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      a[i][j][0] = V->e[i][j][0];
      a[i][j][1] = V->e[i][j][1];
      v[i][j][0] = (i == j);
      v[i][j][1] = 0;
    }
  for (sweep = 0; sweep < 40; sweep++) {
    double off = 0;
    for (p = 0; p < 2; p++)
      for (q = p + 1; q < 3; q++) {
        double app = 0, aqq = 0, gr = 0, gi = 0, g, zeta, t, c, s, er, ei;
        for (i = 0; i < 3; i++) {
          app += a[i][p][0] * a[i][p][0] + a[i][p][1] * a[i][p][1];
          aqq += a[i][q][0] * a[i][q][0] + a[i][q][1] * a[i][q][1];
          gr += a[i][p][0] * a[i][q][0] + a[i][p][1] * a[i][q][1]; /* <a_p, a_q> */
          gi += a[i][p][0] * a[i][q][1] - a[i][p][1] * a[i][q][0];
        }
        g = sqrt(gr * gr + gi * gi);
        if (g <= 1e-300 || g <= 1e-17 * sqrt(app * aqq)) continue;
        off = fmax(off, g / sqrt(app * aqq));
        er = gr / g; /* phase of <a_p, a_q> */
        ei = gi / g;
        zeta = (aqq - app) / (2.0 * g);
        t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        c = 1.0 / sqrt(1.0 + t * t);
        s = c * t;
        /* columns p, q of a and v:  p' = c p - s conj(e) q ,  q' = s e p + c q */
        for (i = 0; i < 3; i++) {
          double (*m)[3][2];
          int which;
          for (which = 0; which < 2; which++) {
            m = which ? v : a;
            {
              const double pr = m[i][p][0], pi = m[i][p][1], qr = m[i][q][0], qi = m[i][q][1];
              /* conj(e) q = (er - i ei)(qr + i qi) */
              const double cqr = er * qr + ei * qi, cqi = er * qi - ei * qr;
              /* e p = (er + i ei)(pr + i pi) */
              const double epr = er * pr - ei * pi, epi = er * pi + ei * pr;
              m[i][p][0] = c * pr - s * cqr;
              m[i][p][1] = c * pi - s * cqi;
              m[i][q][0] = s * epr + c * qr;
              m[i][q][1] = s * epi + c * qi;
            }
          }
        }
      }
    if (off < 1e-15) break;
  }
  /* normalise the columns: U = A J S^-1 */
  for (j = 0; j < 3; j++) {
    double nrm = 0;
    for (i = 0; i < 3; i++) nrm += a[i][j][0] * a[i][j][0] + a[i][j][1] * a[i][j][1];
    nrm = sqrt(nrm);
    for (i = 0; i < 3; i++) {
      a[i][j][0] = nrm > 0 ? a[i][j][0] / nrm : 0;
      a[i][j][1] = nrm > 0 ? a[i][j][1] / nrm : 0;
    }
  }
  /* W = U J^+ */
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      double re = 0, im = 0;
      int k;
      for (k = 0; k < 3; k++) {
        re += a[i][k][0] * v[j][k][0] + a[i][k][1] * v[j][k][1];
        im += a[i][k][1] * v[j][k][0] - a[i][k][0] * v[j][k][1];
      }
      W->e[i][j][0] = re;
      W->e[i][j][1] = im;
    }
}

/* generic_ks/su3_mat_op.c:828-1205 (double branch).  Returns 1 if the SVD branch was taken. */
static int unitarize_one(const mat *V, mat *W, int allow_svd, double svd_rel, double svd_abs) {
  mat Q, Q2;
  double q3d[3], c0, c1, c2, S, R, g0, g1, g2, det_check = 0;
  int i, j, k;
  if (allow_svd) { /* |det V|^2, :849-880 */
    const double (*e)[3][2] = V->e;
    const double a1r = e[1][1][0] * e[2][2][0] - e[1][1][1] * e[2][2][1] - e[1][2][0] * e[2][1][0] + e[1][2][1] * e[2][1][1];
    const double a1i = e[1][1][0] * e[2][2][1] + e[1][1][1] * e[2][2][0] - e[1][2][0] * e[2][1][1] - e[1][2][1] * e[2][1][0];
    const double a2r = e[1][0][0] * e[2][2][0] - e[1][0][1] * e[2][2][1] - e[1][2][0] * e[2][0][0] + e[1][2][1] * e[2][0][1];
    const double a2i = e[1][0][0] * e[2][2][1] + e[1][0][1] * e[2][2][0] - e[1][2][0] * e[2][0][1] - e[1][2][1] * e[2][0][0];
    const double a3r = e[1][0][0] * e[2][1][0] - e[1][0][1] * e[2][1][1] - e[1][1][0] * e[2][0][0] + e[1][1][1] * e[2][0][1];
    const double a3i = e[1][0][0] * e[2][1][1] + e[1][0][1] * e[2][1][0] - e[1][1][0] * e[2][0][1] - e[1][1][1] * e[2][0][0];
    const double dr = e[0][0][0] * a1r - e[0][0][1] * a1i - e[0][1][0] * a2r + e[0][1][1] * a2i + e[0][2][0] * a3r - e[0][2][1] * a3i;
    const double di = e[0][0][1] * a1r + e[0][0][0] * a1i - e[0][1][1] * a2r - e[0][1][0] * a2i + e[0][2][1] * a3r + e[0][2][0] * a3i;
    det_check = dr * dr + di * di;
  }
  mult_an(V, V, &Q);   /* :913-949 */
  mult_nn(&Q, &Q, &Q2); /* :953-960 */
  for (i = 0; i < 3; i++) { /* real part of the diagonal of Q^3, :965-972 */
    double re = 0;
    for (k = 0; k < 3; k++) re += Q2.e[i][k][0] * Q.e[k][i][0] - Q2.e[i][k][1] * Q.e[k][i][1];
    q3d[i] = re;
  }
  c0 = Q.e[0][0][0] + Q.e[1][1][0] + Q.e[2][2][0];
  c1 = (Q2.e[0][0][0] + Q2.e[1][1][0] + Q2.e[2][2][0]) / 2;
  c2 = (q3d[0] + q3d[1] + q3d[2]) / 3;
  S = c1 / 3 - c0 * (c0 / 18);
  if (fabs(S) < KSL_EPS) { /* :985-993 */
    g0 = g1 = g2 = c0 / 3;
  } else {
    double S3, RoS, theta, theta3;
    const double pi23 = 6.28318530717958647692528676656 / 3;
    R = c2 / 2 - c0 * (c1 / 3) + c0 * c0 * (c0 / 27);
    S = sqrt(S);
    S3 = S * S * S;
    RoS = R / S3;
    if (!(fabs(RoS) < 1.0)) theta = (R > 0) ? 0.0 : 3.14159265358979323846264338328;
    else theta = acos(RoS);
    theta3 = theta / 3;
    g0 = c0 / 3 + 2 * S * cos(theta3);
    g1 = c0 / 3 + 2 * S * cos(theta3 + pi23);
    g2 = c0 / 3 + 2 * S * cos(theta3 + 2 * pi23);
  }
  if (allow_svd) { /* :1039-1053 */
    int svd = 0;
    if (det_check != 0 && fabs(det_check - g0 * g1 * g2) / fabs(det_check) > svd_rel) svd = 1;
    if (det_check < svd_abs) svd = 1;
    if (svd) {
      polar_by_svd(V, W);
      return 1;
    }
  }
  { /* :1131-1196 */
    const double g0sq = sqrt(g0), g1sq = sqrt(g1), g2sq = sqrt(g2);
    double us = g1sq + g2sq, ws = g1sq * g2sq, vs = g0sq * us + ws, denom, f0, f1, f2;
    mat S2;
    us += g0sq;
    ws *= g0sq;
    denom = ws * (us * vs - ws);
    f0 = (us * vs * vs - ws * (us * us + vs)) / denom;
    f1 = (2 * us * vs - ws - us * us * us) / denom;
    f2 = us / denom;
    for (i = 0; i < 3; i++)
      for (j = 0; j < 3; j++) {
        S2.e[i][j][0] = f1 * Q.e[i][j][0] + f2 * Q2.e[i][j][0] + (i == j ? f0 : 0.0);
        S2.e[i][j][1] = f1 * Q.e[i][j][1] + f2 * Q2.e[i][j][1];
      }
    mult_nn(V, &S2, W);
  }
  return 0;
}

/* W = U(3) projection of nlinks matrices; returns how many took the SVD branch */
long ksl_unitarize(const double *V, double *W, long nlinks, int allow_svd, double svd_rel, double svd_abs) {
  long k, nsvd = 0;
  for (k = 0; k < nlinks; k++) nsvd += unitarize_one((const mat *)V + k, (mat *)W + k, allow_svd, svd_rel, svd_abs);
  return nsvd;
}

/* U -> V (fat7) -> W (U(3)) -> fat, lng.  coeffs1/coeffs2: the six coefficients of level 1 / 2.
   Any of V, W, fat, lng may be NULL.  Returns the SVD count. */
long ksl_hisq_links(const int *n, const double *coeffs1, const double *coeffs2, const double *links, double *V,
                    double *W, double *fat, double *lng, int allow_svd, double svd_rel, double svd_abs) {
  const long vol = (long)n[0] * n[1] * n[2] * n[3];
  double *v = V ? V : (double *)malloc(sizeof(mat) * 4 * vol);
  double *w = W ? W : (double *)malloc(sizeof(mat) * 4 * vol);
  double *f = fat ? fat : (double *)malloc(sizeof(mat) * 4 * vol);
  long nsvd;
  ksl_smear(n, coeffs1, links, v, NULL);
  nsvd = ksl_unitarize(v, w, 4 * vol, allow_svd, svd_rel, svd_abs);
  ksl_smear(n, coeffs2, w, f, lng);
  if (!V) free(v);
  if (!W) free(w);
  if (!fat) free(f);
  return nsvd;
}
