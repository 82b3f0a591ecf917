/* oracle/ks_force_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C, double precision) of the reference's HISQ fermion force (SURVEY.md
 * section 8 row f2): the checker of milc_qcd_b200/csrc/force.cuh + fermion_force.cu.
 *
 * PARITY PINNED: tests/test_oracle.py checks ksf_hisq_force against the committed output of the
 * reference's own eo_fermion_force_multi (tests/golden/ref_hisq_force.npz, generated from
 * oracle/_ref by tests/golden/make_golden_force.py), live against oracle/_ref when present, and
 * against a finite-difference derivative of the action.
 *
 * What the reference computes (generic_ks/fermion_force_hisq_multi.c:1183-1476, the wrapper_mx
 * path): for S = sum_j res_j |D_oe[U] X_j|^2 with the HISQ chain U -> V (fat7) -> W (U(3)) ->
 * (fat, long) inside D, the momentum update  mom_mu(x) += eps * A_mu(x),  A traceless
 * anti-Hermitian with  dS/dt = -Re tr(i T A)  for  U_mu(x) -> exp(i t T) U_mu(x).
 * The reference walks sorted path tables (:1638-1874) and contracts the derivative of the
 * projection as a rank-4 tensor (u3_unit_der_analytic, generic_ks/su3_mat_op.c).  The
 * restatement is the same derivative organised as reverse-mode differentiation of the forward
 * chain of ks_links_oracle.c -- the organisation the CUDA kernels will use:
 *   1. outer products  G_fat_mu(x) = +-2 res_j Z_j(x) Z_j(x+mu)^+,  G_lng likewise with x+3mu
 *      (Z = X on even sites, D X on odd sites; sign + on odd x)   cf. outer_product_append, :2009-2154
 *   2. level-2 smearing and Naik product backwards: every staple pass of load_fatlinks_cpu gives
 *      six gradient contributions (three per staple)               cf. :1638-1874 with the p2 table
 *   3. U(3) projection backwards: W = V Q^-1/2, Q = V^+ V, with the exact derivative of the
 *      matrix function from the eigen-decomposition of Q (Daleckii-Krein)   cf. :1877-2006
 *   4. level-1 (fat7) smearing backwards                            cf. :1638-1874 with the p1 table
 *   5. A = -TA(U G_U^+), packed as anti_hermitmat                    cf. :1433-1470
 * "G_M" is the gradient matrix defined by dS = Re tr(G_M^+ dM).
 * HISQ_FORCE_FILTER (su3_mat_op.c:1680-1734, 5e-5 in ks_imp_rhmc's build) is restated: on a link whose
 * smallest eigenvalue of Q is below the filter, step 3 differentiates V (Q + filter)^-1/2 instead.
 * That function is not unitary, so on those links the RADIAL part of the level-2 gradient matters,
 * and with it the form in which the Lepage term is differentiated (see smear_bwd); with both
 * the restatement meets the reference to 1e-10 on links rough enough to trip the filter
 * (tests/test_oracle.py, golden ref_hisq_force_rough.npz).
 * Several Naik epsilons: ksf_hisq_force_naik, pinned live against the reference (tests/test_oracle.py) and on
 * tests/golden/ref_hisq_force_naik.npz.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double e[3][3][2]; } mat;

/* from ks_links_oracle.c */
void ksl_smear(const int *n, const double *coeffs, const double *links, double *fat, double *lng);
long ksl_unitarize(const double *V, double *W, long nlinks, int allow_svd, double svd_rel, double svd_abs);

static long f_node_index(const int *n, int x, int y, int z, int t) {
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  long lex = x + (long)n[0] * (y + (long)n[1] * (z + (long)n[2] * t));
  return (((x + y + z + t) & 1) == 0) ? lex / 2 : (lex + vol) / 2;
}
static int *f_build_nb(const int *n, int d) { /* d < 4: +1 in d ; d >= 4: -1 in d-4 */
  long vol = (long)n[0] * n[1] * n[2] * n[3];
  int *nb = (int *)malloc(sizeof(int) * vol);
  int x[4];
  for (x[3] = 0; x[3] < n[3]; x[3]++)
    for (x[2] = 0; x[2] < n[2]; x[2]++)
      for (x[1] = 0; x[1] < n[1]; x[1]++)
        for (x[0] = 0; x[0] < n[0]; x[0]++) {
          int y[4] = {x[0], x[1], x[2], x[3]}, mu = d & 3;
          y[mu] = (x[mu] + (d < 4 ? 1 : -1) + n[mu]) % n[mu];
          nb[f_node_index(n, x[0], x[1], x[2], x[3])] = (int)f_node_index(n, y[0], y[1], y[2], y[3]);
        }
  return nb;
}

static void nn(const mat *a, const mat *b, mat *c) {
  int i, j, k;
  mat r;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (k = 0; k < 3; k++) {
        re += a->e[i][k][0] * b->e[k][j][0] - a->e[i][k][1] * b->e[k][j][1];
        im += a->e[i][k][0] * b->e[k][j][1] + a->e[i][k][1] * b->e[k][j][0];
      }
      r.e[i][j][0] = re;
      r.e[i][j][1] = im;
    }
  *c = r;
}
static void adj(const mat *a, mat *c) {
  int i, j;
  mat r;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      r.e[i][j][0] = a->e[j][i][0];
      r.e[i][j][1] = -a->e[j][i][1];
    }
  *c = r;
}
static void na(const mat *a, const mat *b, mat *c) { mat t; adj(b, &t); nn(a, &t, c); }
static void an(const mat *a, const mat *b, mat *c) { mat t; adj(a, &t); nn(&t, b, c); }
static void axpy(mat *a, double s, const mat *b) {
  int i, j, r;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++)
      for (r = 0; r < 2; r++) a->e[i][j][r] += s * b->e[i][j][r];
}

/* forward staple, generic/general_staple.c:41-123 (no accumulation into a fat link here);
   part: 1 = the upper staple only, 2 = the lower one only, 3 = both */
static void staple_fwd(long vol, int *const *nb, mat *out, int mu, int nu, const mat *link, int stride, const mat *links, int part) {
  long i;
  for (i = 0; i < vol; i++) {
    const long y = nb[4 + nu][i];
    mat t1, up, low;
    memset(&out[i], 0, sizeof(mat));
    if (part & 1) {
      na(&link[(long)stride * nb[nu][i]], &links[4l * nb[mu][i] + nu], &t1);
      nn(&links[4 * i + nu], &t1, &up);
      axpy(&out[i], 1.0, &up);
    }
    if (part & 2) {
      an(&links[4 * y + nu], &link[(long)stride * y], &t1);
      nn(&t1, &links[4l * nb[mu][y] + nu], &low);
      axpy(&out[i], 1.0, &low);
    }
  }
}

/* gradient of one staple pass: H = gradient w.r.t. the staple field; adds to g_link (gradient
   w.r.t. the field standing in for the mu link, same stride) and to g_links (gauge links).
     upper(x) = A B C^+ : A = U_nu(x), B = link(x+nu), C = U_nu(x+mu)
     lower(x) = D^+ E F : D = U_nu(y), E = link(y),  F = U_nu(y+mu),  y = x-nu            */
static void staple_bwd(long vol, int *const *nb, const mat *H, int mu, int nu, const mat *link, int stride,
                       const mat *links, mat *g_link, mat *g_links, int part) {
  long x;
  for (x = 0; x < vol; x++) {
    const long xpn = nb[nu][x], xpm = nb[mu][x], y = nb[4 + nu][x], ypm = nb[mu][y];
    const mat *A = &links[4 * x + nu], *B = &link[(long)stride * xpn], *C = &links[4 * xpm + nu];
    const mat *D = &links[4 * y + nu], *E = &link[(long)stride * y], *F = &links[4 * ypm + nu];
    const mat *h = &H[x];
    mat t1, t2;
    if (part & 1) { /* upper */
      na(C, B, &t1); nn(h, &t1, &t2); axpy(&g_links[4 * x + nu], 1.0, &t2);          /* G_A += H C B^+ */
      an(A, h, &t1); nn(&t1, C, &t2); axpy(&g_link[(long)stride * xpn], 1.0, &t2);   /* G_B += A^+ H C */
      an(h, A, &t1); nn(&t1, B, &t2); axpy(&g_links[4 * xpm + nu], 1.0, &t2);        /* G_C += H^+ A B */
    }
    if (part & 2) { /* lower */
      nn(D, h, &t1); na(&t1, F, &t2); axpy(&g_link[(long)stride * y], 1.0, &t2);     /* G_E += D H F^+ */
      an(E, D, &t1); nn(&t1, h, &t2); axpy(&g_links[4 * ypm + nu], 1.0, &t2);        /* G_F += E^+ D H */
      nn(E, F, &t1); na(&t1, h, &t2); axpy(&g_links[4 * y + nu], 1.0, &t2);          /* G_D += E F H^+ */
    }
  }
}

/* reverse of ksl_smear: g_fat (and g_lng, may be NULL) -> adds to g_links.
   The Lepage term is differentiated in the form the reference's force walks it (its path table holds the
   straight double staple +nu+nu+mu-nu-nu only, with the one-link coefficient as given: imp_actions/hisq/
   hisq_u3_action.h path_coeff_2, fermion_force_hisq_multi.c:1638-1874), not in the form the fattening computes it
   (staple of the staple in the same direction, whose two back-tracking terms U_nu U_nu^+ U_mu ... are cancelled by
   the "one_link - 6 lepage" coefficient, fermion_links_fn_load_milc.c:146).  The two functions agree for unitary
   links and in every direction tangent to the group, so the force is the same; they differ in the radial part of
   the gradient, which only the filtered links of step 3 can see. */
static void smear_bwd(const int *n, const double *coeffs, const mat *links, const mat *g_fat, const mat *g_lng, mat *g_links) {
  const long vol = (long)n[0] * n[1] * n[2] * n[3];
  const double one_link = coeffs[0], naik = coeffs[1], three = coeffs[2], five = coeffs[3], seven = coeffs[4],
               lepage = coeffs[5];
  int *nb[8];
  int d, dir, nu, rho, sig, part;
  long i;
  mat *st3 = (mat *)malloc(sizeof(mat) * vol), *st5 = (mat *)malloc(sizeof(mat) * vol);
  mat *g3 = (mat *)malloc(sizeof(mat) * vol), *g5 = (mat *)malloc(sizeof(mat) * vol), *h = (mat *)malloc(sizeof(mat) * vol);
  mat *stp = (mat *)malloc(sizeof(mat) * vol), *gp = (mat *)malloc(sizeof(mat) * vol);
  for (d = 0; d < 8; d++) nb[d] = f_build_nb(n, d);
  for (dir = 0; dir < 4; dir++) {
    for (i = 0; i < vol; i++) axpy(&g_links[4 * i + dir], one_link, &g_fat[4 * i + dir]);
    if (three == 0.0 && lepage == 0.0 && five == 0.0) continue;
    for (nu = 0; nu < 4; nu++) {
      if (nu == dir) continue;
      staple_fwd(vol, nb, st3, dir, nu, links + dir, 4, links, 3);
      for (i = 0; i < vol; i++) { /* gradient reaching the 3-staple directly: c3 * G_fat */
        memset(&g3[i], 0, sizeof(mat));
        axpy(&g3[i], three, &g_fat[4 * i + dir]);
      }
      if (lepage != 0.0) /* fat += lepage * (upper staple of the upper 3-staple + lower of the lower) */
        for (part = 1; part <= 2; part++) {
          staple_fwd(vol, nb, stp, dir, nu, links + dir, 4, links, part);
          for (i = 0; i < vol; i++) { memset(&h[i], 0, sizeof(mat)); axpy(&h[i], lepage, &g_fat[4 * i + dir]); memset(&gp[i], 0, sizeof(mat)); }
          staple_bwd(vol, nb, h, dir, nu, stp, 1, links, gp, g_links, part);
          staple_bwd(vol, nb, gp, dir, nu, links + dir, 4, links, g_links + dir, g_links, part);
        }
      for (rho = 0; rho < 4; rho++) {
        if (rho == dir || rho == nu) continue;
        staple_fwd(vol, nb, st5, dir, rho, st3, 1, links, 3);
        for (i = 0; i < vol; i++) { memset(&g5[i], 0, sizeof(mat)); axpy(&g5[i], five, &g_fat[4 * i + dir]); }
        for (sig = 0; sig < 4; sig++) {
          if (sig == dir || sig == nu || sig == rho) continue;
          for (i = 0; i < vol; i++) { memset(&h[i], 0, sizeof(mat)); axpy(&h[i], seven, &g_fat[4 * i + dir]); }
          staple_bwd(vol, nb, h, dir, sig, st5, 1, links, g5, g_links, 3);
        }
        staple_bwd(vol, nb, g5, dir, rho, st3, 1, links, g3, g_links, 3);
      }
      staple_bwd(vol, nb, g3, dir, nu, links + dir, 4, links, g_links + dir, g_links, 3);
    }
  }
  if (g_lng) /* lng(x) = naik U(x) U(x+mu) U(x+2mu) */
    for (dir = 0; dir < 4; dir++)
      for (i = 0; i < vol; i++) {
        const long i1 = nb[dir][i], i2 = nb[dir][i1];
        const mat *a = &links[4 * i + dir], *b = &links[4 * i1 + dir], *c = &links[4 * i2 + dir], *g = &g_lng[4 * i + dir];
        mat t1, t2;
        nn(b, c, &t1); na(g, &t1, &t2); axpy(&g_links[4 * i + dir], naik, &t2);      /* G_a += naik G (b c)^+ */
        an(a, g, &t1); na(&t1, c, &t2); axpy(&g_links[4 * i1 + dir], naik, &t2);    /* G_b += naik a^+ G c^+ */
        nn(a, b, &t1); an(&t1, g, &t2); axpy(&g_links[4 * i2 + dir], naik, &t2);    /* G_c += naik (a b)^+ G */
      }
  for (d = 0; d < 8; d++) free(nb[d]);
  free(st3); free(st5); free(g3); free(g5); free(h); free(stp); free(gp);
}

/* Hermitian 3x3 eigen-decomposition by cyclic Jacobi: Q = sum_k g_k |v_k><v_k|, v_k = column k */
static void herm_eig(const mat *Q, double g[3], mat *vecs) {
  double a[3][3][2], v[3][3][2];
  int i, j, p, q, sweep;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      a[i][j][0] = Q->e[i][j][0]; a[i][j][1] = Q->e[i][j][1];
      v[i][j][0] = (i == j); v[i][j][1] = 0;
    }
  for (sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (p = 0; p < 2; p++)
      for (q = p + 1; q < 3; q++) off += a[p][q][0] * a[p][q][0] + a[p][q][1] * a[p][q][1];
    if (off < 1e-60) break;
    for (p = 0; p < 2; p++)
      for (q = p + 1; q < 3; q++) {
        const double ar = a[p][q][0], ai = a[p][q][1], ab = sqrt(ar * ar + ai * ai);
        double er, ei, theta, t, c, s;
        if (ab < 1e-300) continue;
        er = ar / ab; ei = ai / ab; /* a_pq = ab e^{i phi} */
        theta = (a[q][q][0] - a[p][p][0]) / (2.0 * ab);
        t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
        c = 1.0 / sqrt(1.0 + t * t); s = c * t;
        /* unitary rotation R: columns p,q:  p' = c p - s e^{-i phi} q ,  q' = s e^{i phi} p + c q ; A <- R^+ A R */
        for (i = 0; i < 3; i++) { /* columns of a and v */
          int w;
          for (w = 0; w < 2; w++) {
            double (*m)[3][2] = w ? v : a;
            const double pr = m[i][p][0], pi = m[i][p][1], qr = m[i][q][0], qi = m[i][q][1];
            const double cqr = er * qr + ei * qi, cqi = er * qi - ei * qr;   /* e^{-i phi} q */
            const double epr = er * pr - ei * pi, epi = er * pi + ei * pr;   /* e^{+i phi} p */
            m[i][p][0] = c * pr - s * cqr; m[i][p][1] = c * pi - s * cqi;
            m[i][q][0] = s * epr + c * qr; m[i][q][1] = s * epi + c * qi;
          }
        }
        for (j = 0; j < 3; j++) { /* rows of a: row p' = c row p - s e^{+i phi} row q ; row q' = s e^{-i phi} row p + c row q */
          const double pr = a[p][j][0], pi = a[p][j][1], qr = a[q][j][0], qi = a[q][j][1];
          const double eqr = er * qr - ei * qi, eqi = er * qi + ei * qr;     /* e^{+i phi} q */
          const double cpr = er * pr + ei * pi, cpi = er * pi - ei * pr;     /* e^{-i phi} p */
          a[p][j][0] = c * pr - s * eqr; a[p][j][1] = c * pi - s * eqi;
          a[q][j][0] = s * cpr + c * qr; a[q][j][1] = s * cpi + c * qi;
        }
      }
  }
  for (i = 0; i < 3; i++) g[i] = a[i][i][0];
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) { vecs->e[i][j][0] = v[i][j][0]; vecs->e[i][j][1] = v[i][j][1]; }
}

/* reverse of W = V Q^-1/2, Q = V^+ V:  G_V = G_W Q^-1/2 + V (G_Q + G_Q^+),
   G_Q = sum_ij phi_ij P_i R P_j,  R = V^+ G_W,  phi_ij = (g_i^-1/2 - g_j^-1/2)/(g_i - g_j)  (-g^-3/2 / 2 on the diagonal) */
static void unitarize_bwd(const mat *V, const mat *GW, mat *GV, double filter) {
  mat Q, E, Ed, R, Rt, Gq, Gqd, S, t1, t2;
  double g[3], phi[3][3], gmin;
  int i, j;
  an(V, V, &Q);
  herm_eig(&Q, g, &E);
  adj(&E, &Ed);
  /* HISQ_FORCE_FILTER (su3_mat_op.c:1680-1734): when the smallest eigenvalue of Q is below the filter, the
     reference adds the filter to all three and to Q's diagonal, i.e. it differentiates V (V^+ V + filter)^-1/2 */
  gmin = g[0] < g[1] ? g[0] : g[1];
  if (g[2] < gmin) gmin = g[2];
  if (filter > 0 && gmin < filter)
    for (i = 0; i < 3; i++) g[i] += filter;
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) {
      const double si = 1.0 / sqrt(g[i]), sj = 1.0 / sqrt(g[j]);
      if (i == j || fabs(g[i] - g[j]) < 1e-9 * (g[i] + g[j])) phi[i][j] = -0.5 * si * si * si;   /* = -1/(si^-1 sj^-1 (si^-1 + sj^-1)) at g_i = g_j */
      else phi[i][j] = -1.0 / ((1.0 / si) * (1.0 / sj) * (1.0 / si + 1.0 / sj));               /* exact: (si - sj)/(g_i - g_j) */
    }
  /* Q^-1/2 = E diag(g^-1/2) E^+ */
  memset(&S, 0, sizeof(S));
  for (i = 0; i < 3; i++) { S.e[i][i][0] = 1.0 / sqrt(g[i]); }
  nn(&E, &S, &t1); nn(&t1, &Ed, &S);
  an(V, GW, &R);
  /* rotate into the eigenbasis, scale, rotate back */
  nn(&Ed, &R, &t1); nn(&t1, &E, &Rt);
  for (i = 0; i < 3; i++)
    for (j = 0; j < 3; j++) { Rt.e[i][j][0] *= phi[i][j]; Rt.e[i][j][1] *= phi[i][j]; }
  nn(&E, &Rt, &t1); nn(&t1, &Ed, &Gq);
  adj(&Gq, &Gqd);
  axpy(&Gq, 1.0, &Gqd);
  nn(GW, &S, &t1);
  nn(V, &Gq, &t2);
  *GV = t1;
  axpy(GV, 1.0, &t2);
}

/* one link of step 3, exported for the tests: G_V from V (9 complex) and G_W */
void ksf_unitarize_bwd(const double *V, const double *GW, double *GV, double filter) {
  unitarize_bwd((const mat *)V, (const mat *)GW, (mat *)GV, filter);
}

/* the reference build's HISQ_FORCE_FILTER (oracle/build_ref.sh, Make_template default 5.0e-5); 0 = unfiltered */
double ksf_force_filter = 5.0e-5;
void ksf_set_force_filter(double f) { ksf_force_filter = f; }

/* outer products of nterms terms, added to gfat (1-hop) and glng (3-hop):  +-2 res_j Z_j(x) Z_j(x+mu)^+ */
static void oprods(long vol, int *const *nbp, const double *multi_x, const double *residues, int nterms, mat *gfat, mat *glng) {
  int dir, j, a, b;
  long i;
  for (j = 0; j < nterms; j++) {
    const double *Z = multi_x + (size_t)j * vol * 6;
    for (i = 0; i < vol; i++) {
      const double sgn = (i >= vol / 2 ? 2.0 : -2.0) * residues[j];   /* odd sites come second in MILC's order */
      for (dir = 0; dir < 4; dir++) {
        const long i1 = nbp[dir][i], i3 = nbp[dir][nbp[dir][i1]];
        for (a = 0; a < 3; a++)
          for (b = 0; b < 3; b++) {
            const double zr = Z[6 * i + 2 * a], zi = Z[6 * i + 2 * a + 1];
            const double pr = Z[6 * i1 + 2 * b], pi = Z[6 * i1 + 2 * b + 1];
            const double qr = Z[6 * i3 + 2 * b], qi = Z[6 * i3 + 2 * b + 1];
            gfat[4 * i + dir].e[a][b][0] += sgn * (zr * pr + zi * pi);   /* z conj(p) */
            gfat[4 * i + dir].e[a][b][1] += sgn * (zi * pr - zr * pi);
            glng[4 * i + dir].e[a][b][0] += sgn * (zr * qr + zi * qi);
            glng[4 * i + dir].e[a][b][1] += sgn * (zi * qr - zr * qi);
          }
      }
    }
  }
}

/* multi_x: nterms fields of V colour vectors [6 doubles per site] (even sites X_j, odd sites D X_j);
   mom out: anti_hermitmat[4*V] as 10 doubles {m01.re, m01.im, m02.re, m02.im, m12.re, m12.im, m00im, m11im, m22im, 0}.
   Several Naik epsilons (fermion_force_hisq_multi.c:1285-1375; links: fermion_links_hisq_load_milc.c:573-636):
   the terms come in n_naiks classes of n_orders[k] terms; class k was solved with
       fat_k = fat_0 + eps_naik[k] * coeffs3[0] * W ,   lng_k = lng_0 + eps_naik[k] * coeffs3[1] * W W W
   (eps_naik[0] = 0; coeffs3 = the one-link and Naik coefficients of the reference's third path table), so every
   term goes through the level-2 smearing and the terms of class k >= 1 add
       G_W += eps_naik[k] * ( coeffs3[0] * G_fat^(k) + Naik product backwards of coeffs3[1] * G_lng^(k) ). */
void ksf_hisq_force_naik(const int *n, const double *coeffs1, const double *coeffs2, const double *coeffs3, const double *links_,
                         const double *multi_x, const double *residues, int n_naiks, const int *n_orders,
                         const double *eps_naik, double eps, double *mom) {
  const long vol = (long)n[0] * n[1] * n[2] * n[3];
  const mat *U = (const mat *)links_;
  mat *V = (mat *)malloc(sizeof(mat) * 4 * vol), *W = (mat *)malloc(sizeof(mat) * 4 * vol);
  mat *gfat = (mat *)calloc(4 * vol, sizeof(mat)), *glng = (mat *)calloc(4 * vol, sizeof(mat));
  mat *gW = (mat *)calloc(4 * vol, sizeof(mat)), *gV = (mat *)malloc(sizeof(mat) * 4 * vol), *gU = (mat *)calloc(4 * vol, sizeof(mat));
  int *nbp[4];
  int dir, k, nterms = 0, shift;
  long i;
  for (dir = 0; dir < 4; dir++) nbp[dir] = f_build_nb(n, dir);
  for (k = 0; k < n_naiks; k++) nterms += n_orders[k];
  ksl_smear(n, coeffs1, links_, (double *)V, NULL);
  ksl_unitarize((const double *)V, (double *)W, 4 * vol, 0, 0.0, 0.0);
  /* 1. outer products */
  oprods(vol, nbp, multi_x, residues, nterms, gfat, glng);
  /* 2.-4. the chain backwards */
  smear_bwd(n, coeffs2, W, gfat, glng, gW);
  shift = n_orders[0];
  for (k = 1; k < n_naiks; k++) { /* the one-link + Naik table of class k, weighted with its epsilon */
    const double c3[6] = {eps_naik[k] * coeffs3[0], eps_naik[k] * coeffs3[1], 0, 0, 0, 0};
    memset(gfat, 0, sizeof(mat) * 4 * vol);
    memset(glng, 0, sizeof(mat) * 4 * vol);
    oprods(vol, nbp, multi_x + (size_t)shift * vol * 6, residues + shift, n_orders[k], gfat, glng);
    smear_bwd(n, c3, W, gfat, glng, gW);
    shift += n_orders[k];
  }
  for (i = 0; i < 4 * vol; i++) unitarize_bwd(&V[i], &gW[i], &gV[i], ksf_force_filter);
  smear_bwd(n, coeffs1, U, gV, NULL, gU);
  /* 5. A = -eps TA(U G_U^+) */
  for (i = 0; i < 4 * vol; i++) {
    mat M;
    double tr;
    double *m = mom + 10 * i;
    na(&U[i], &gU[i], &M);
    /* anti-Hermitian part (M - M^+)/2 */
    tr = (M.e[0][0][1] + M.e[1][1][1] + M.e[2][2][1]) / 3.0;
    m[0] = -eps * 0.5 * (M.e[0][1][0] - M.e[1][0][0]);
    m[1] = -eps * 0.5 * (M.e[0][1][1] + M.e[1][0][1]);
    m[2] = -eps * 0.5 * (M.e[0][2][0] - M.e[2][0][0]);
    m[3] = -eps * 0.5 * (M.e[0][2][1] + M.e[2][0][1]);
    m[4] = -eps * 0.5 * (M.e[1][2][0] - M.e[2][1][0]);
    m[5] = -eps * 0.5 * (M.e[1][2][1] + M.e[2][1][1]);
    m[6] = -eps * (M.e[0][0][1] - tr);
    m[7] = -eps * (M.e[1][1][1] - tr);
    m[8] = -eps * (M.e[2][2][1] - tr);
    m[9] = 0.0;
  }
  for (dir = 0; dir < 4; dir++) free(nbp[dir]);
  free(V); free(W); free(gfat); free(glng); free(gW); free(gV); free(gU);
}

void ksf_hisq_force(const int *n, const double *coeffs1, const double *coeffs2, const double *links_, const double *multi_x,
                    const double *residues, int nterms, double eps, double *mom) {
  const double zero = 0.0, c3[2] = {0.0, 0.0};
  ksf_hisq_force_naik(n, coeffs1, coeffs2, c3, links_, multi_x, residues, 1, &nterms, &zero, eps, mom);
}
