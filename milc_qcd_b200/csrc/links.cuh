// links.cuh -- HISQ / asqtad fermion-link construction on the device.
//
// What the reference computes on the CPU for every new gauge field (twice per RHMC step):
//   level 1   V = fat7(U)                       load_fatlinks_cpu, generic_ks/fermion_links_fn_load_milc.c:120-275
//   project   W = V (V^+ V)^-1/2  (U(3))         u3_unitarize_analytic, generic_ks/su3_mat_op.c:828-1205
//   level 2   fat = asqtad-like smear of W,      load_fatlinks_cpu again (+ Lepage term)
//             lng = c_naik W W W                 load_lnglinks, :45-107
// (generic_ks/fermion_links_hisq_load_milc.c:531-586).  The seam MILC already has for it is
// qudaLoadUnitarizedLink / qudaLoadKSLink (generic_ks/fermion_links_fn_load_gpu.c:18-123).
//
// Smearing follows the reference's recursion (ASQ_OPTIMIZED_FATTENING): per direction mu the
// 3-staples in nu, the Lepage staple of each 3-staple, the 5-staples (staples of the 3-staples
// in rho), the 7-staples (staples of the 5-staples in sigma), each added to fat_mu with its
// coefficient -- 18 staple passes per direction (generic/general_staple.c:41-123).  One thread
// per site computes the upper and the lower staple in place (the reference builds the lower one
// at x-nu and gathers it; same products, same order).
//
// Layout: "matrix fields" are 9 planes of double2 with plane stride fstride; site index
// f = parity*Vh + cb, i.e. MILC's own site order (even sites, then odd sites), so host links
// su3_matrix[4*i+dir] transpose straight in (pack_link_kernel per parity half).  A link field is
// four matrix fields, plane (mu*9 + e).  Single GPU; double precision throughout, as the reference
// does the projection in double whatever MILC_PRECISION is.
#pragma once
#include "common.cuh"
#include "synth.cuh"

namespace b200ks {

struct M3 { double2 e[9]; };

__device__ __forceinline__ M3 m3_load(const double2 *p, size_t fstride, int f) {
  M3 m;
#pragma unroll
  for (int k = 0; k < 9; k++) m.e[k] = p[(size_t)k * fstride + f];
  return m;
}
__device__ __forceinline__ void m3_store(double2 *p, size_t fstride, int f, const M3 &m) {
#pragma unroll
  for (int k = 0; k < 9; k++) p[(size_t)k * fstride + f] = m.e[k];
}
// c = a b
__device__ __forceinline__ M3 m3_nn(const M3 &a, const M3 &b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double2 x = a.e[3 * i + k], y = b.e[3 * k + j];
        re = fma(x.x, y.x, re);
        re = fma(-x.y, y.y, re);
        im = fma(x.x, y.y, im);
        im = fma(x.y, y.x, im);
      }
      c.e[3 * i + j] = make_double2(re, im);
    }
  return c;
}
// c = a b^dagger
__device__ __forceinline__ M3 m3_na(const M3 &a, const M3 &b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double2 x = a.e[3 * i + k], y = b.e[3 * j + k];
        re = fma(x.x, y.x, re);
        re = fma(x.y, y.y, re);
        im = fma(x.y, y.x, im);
        im = fma(-x.x, y.y, im);
      }
      c.e[3 * i + j] = make_double2(re, im);
    }
  return c;
}
// c = a^dagger b
__device__ __forceinline__ M3 m3_an(const M3 &a, const M3 &b) {
  M3 c;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const double2 x = a.e[3 * k + i], y = b.e[3 * k + j];
        re = fma(x.x, y.x, re);
        re = fma(x.y, y.y, re);
        im = fma(x.x, y.y, im);
        im = fma(-x.y, y.x, im);
      }
      c.e[3 * i + j] = make_double2(re, im);
    }
  return c;
}
__device__ __forceinline__ void m3_axpy(M3 &a, double s, const M3 &b) {
#pragma unroll
  for (int k = 0; k < 9; k++) {
    a.e[k].x = fma(s, b.e[k].x, a.e[k].x);
    a.e[k].y = fma(s, b.e[k].y, a.e[k].y);
  }
}

// full-lattice site index of the neighbour of f at +-1 in direction mu (periodic)
__device__ __forceinline__ int full_neighbor(const Geom &g, int f, int mu, int sign) {
  const int par = f >= g.Vh ? 1 : 0;
  const int cb = f - par * g.Vh;
  const Coord c = site_coord(g, cb, par);
  int n;
  if (sign > 0) {
    n = mu == 0 ? neighbor<0>(g, cb, c, 1) : mu == 1 ? neighbor<1>(g, cb, c, 1) : mu == 2 ? neighbor<2>(g, cb, c, 1) : neighbor<3>(g, cb, c, 1);
  } else {
    n = mu == 0 ? neighbor<0>(g, cb, c, -1) : mu == 1 ? neighbor<1>(g, cb, c, -1) : mu == 2 ? neighbor<2>(g, cb, c, -1) : neighbor<3>(g, cb, c, -1);
  }
  return (par ^ 1) * g.Vh + n;
}

// fat_mu = c1 * U_mu for the four directions      (fermion_links_fn_load_milc.c:214-222)
__global__ void __launch_bounds__(kBlock)
onelink_kernel(double2 *fat, const double2 *links, double c1, size_t fstride, int nsites) {
  const int f = blockIdx.x * kBlock + threadIdx.x;
  if (f >= nsites) return;
#pragma unroll 4
  for (int m = 0; m < 36; m++) {
    const double2 u = links[(size_t)m * fstride + f];
    fat[(size_t)m * fstride + f] = make_double2(c1 * u.x, c1 * u.y);
  }
}

// One staple pass, generic/general_staple.c:41-123.  `link` is the matrix field standing in for
// the mu link (the gauge link itself for the 3-staple, a staple field for the others); `links`
// the gauge field whose nu links close the staple.
//   upper(x) = U_nu(x) link(x+nu) U_nu(x+mu)^+          lower(x) = U_nu(y)^+ link(y) U_nu(y+mu), y = x-nu
// kSave: staple_out = upper + lower and fat_mu += coef*staple_out; else fat_mu += coef*upper, then
// += coef*lower (the reference's two orders of summation).
template <bool kSave>
__global__ void __launch_bounds__(kBlock)
staple_kernel(double2 *staple_out, const double2 *link, const double2 *links, double2 *fat, int mu, int nu, double coef,
              const Geom g, size_t fstride, int nsites, int interleaved) {
  int f = blockIdx.x * kBlock + threadIdx.x;
  if (interleaved) f = interleaved_site(f, g.Vh);   // (common.cuh: both uses of a matrix a few CTAs apart)
  else if (f >= nsites) f = -1;
  if (f < 0) return;
  const double2 *Unu = links + (size_t)nu * 9 * fstride;
  const int f_pnu = full_neighbor(g, f, nu, 1), f_pmu = full_neighbor(g, f, mu, 1);
  const int f_mnu = full_neighbor(g, f, nu, -1), f_mnu_pmu = full_neighbor(g, f_mnu, mu, 1);
  M3 up, low;
  {
    const M3 t1 = m3_na(m3_load(link, fstride, f_pnu), m3_load(Unu, fstride, f_pmu));
    up = m3_nn(m3_load(Unu, fstride, f), t1);
  }
  {
    const M3 t1 = m3_an(m3_load(Unu, fstride, f_mnu), m3_load(link, fstride, f_mnu));
    low = m3_nn(t1, m3_load(Unu, fstride, f_mnu_pmu));
  }
  double2 *fmu = fat + (size_t)mu * 9 * fstride;
  M3 acc = m3_load(fmu, fstride, f);
  if (kSave) {
#pragma unroll
    for (int k = 0; k < 9; k++) {
      up.e[k].x += low.e[k].x;
      up.e[k].y += low.e[k].y;
    }
    m3_store(staple_out, fstride, f, up);
    m3_axpy(acc, coef, up);
  } else {
    m3_axpy(acc, coef, up);
    m3_axpy(acc, coef, low);
  }
  m3_store(fmu, fstride, f, acc);
}

// lng_mu(x) = naik * U_mu(x) U_mu(x+mu) U_mu(x+2mu)    (fermion_links_fn_load_milc.c:45-107)
__global__ void __launch_bounds__(kBlock)
longlink_kernel(double2 *lng, const double2 *links, double naik, const Geom g, size_t fstride, int nsites) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k >= 4 * nsites) return;
  const int mu = k / nsites, f = k - mu * nsites;
  const double2 *U = links + (size_t)mu * 9 * fstride;
  const int f1 = full_neighbor(g, f, mu, 1), f2 = full_neighbor(g, f1, mu, 1);
  const M3 t = m3_nn(m3_nn(m3_load(U, fstride, f), m3_load(U, fstride, f1)), m3_load(U, fstride, f2));
  M3 o;
#pragma unroll
  for (int e = 0; e < 9; e++) o.e[e] = make_double2(naik * t.e[e].x, naik * t.e[e].y);
  m3_store(lng + (size_t)mu * 9 * fstride, fstride, f, o);
}

// ---- U(3) projection ----------------------------------------------------------------------------
// One-sided (Hestenes) Jacobi SVD: rotations from the right make the columns of A orthogonal,
// A J = U S; the unitary polar factor is U J^+ -- what the reference forms from svd3x3's factors
// (su3_mat_op.c:1059-1110) when the Cayley-Hamilton eigenvalues fail the determinant check.
__device__ inline M3 polar_by_svd(const M3 &V) {
  double2 a[3][3], v[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      a[i][j] = V.e[3 * i + j];
      v[i][j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
  for (int sweep = 0; sweep < 40; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double app = 0, aqq = 0, gr = 0, gi = 0;
        for (int i = 0; i < 3; i++) {
          app += a[i][p].x * a[i][p].x + a[i][p].y * a[i][p].y;
          aqq += a[i][q].x * a[i][q].x + a[i][q].y * a[i][q].y;
          gr += a[i][p].x * a[i][q].x + a[i][p].y * a[i][q].y;
          gi += a[i][p].x * a[i][q].y - a[i][p].y * a[i][q].x;
        }
        const double gg = sqrt(gr * gr + gi * gi);
        if (gg <= 1e-300 || gg <= 1e-17 * sqrt(app * aqq)) continue;
        off = fmax(off, gg / sqrt(app * aqq));
        const double er = gr / gg, ei = gi / gg;
        const double zeta = (aqq - app) / (2.0 * gg);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++)
          for (int which = 0; which < 2; which++) {
            double2 &P = which ? v[i][p] : a[i][p];
            double2 &Q = which ? v[i][q] : a[i][q];
            const double pr = P.x, pi = P.y, qr = Q.x, qi = Q.y;
            const double cqr = er * qr + ei * qi, cqi = er * qi - ei * qr;   // conj(e) q
            const double epr = er * pr - ei * pi, epi = er * pi + ei * pr;   // e p
            P = make_double2(c * pr - s * cqr, c * pi - s * cqi);
            Q = make_double2(s * epr + c * qr, s * epi + c * qi);
          }
      }
    if (off < 1e-15) break;
  }
  for (int j = 0; j < 3; j++) {
    double nrm = 0;
    for (int i = 0; i < 3; i++) nrm += a[i][j].x * a[i][j].x + a[i][j].y * a[i][j].y;
    nrm = sqrt(nrm);
    for (int i = 0; i < 3; i++) a[i][j] = nrm > 0 ? make_double2(a[i][j].x / nrm, a[i][j].y / nrm) : make_double2(0.0, 0.0);
  }
  M3 W;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (int k = 0; k < 3; k++) {
        re += a[i][k].x * v[j][k].x + a[i][k].y * v[j][k].y;
        im += a[i][k].y * v[j][k].x - a[i][k].x * v[j][k].y;
      }
      W.e[3 * i + j] = make_double2(re, im);
    }
  return W;
}

struct ReunitParams {
  int allow_svd;        // HISQ_REUNIT_ALLOW_SVD
  double svd_rel;       // HISQ_REUNIT_SVD_REL_ERROR
  double svd_abs;       // HISQ_REUNIT_SVD_ABS_ERROR
};

// u3_unitarize_analytic, generic_ks/su3_mat_op.c:828-1205 (double branch): Q = V^+ V, eigenvalues of
// Q from its characteristic polynomial (Cayley-Hamilton, trigonometric solution), Q^-1/2 =
// f0 + f1 Q + f2 Q^2, W = V Q^-1/2; the SVD branch when |det V|^2 is tiny or disagrees with the
// product of the eigenvalues.  Returns true if the SVD branch was taken.
__device__ inline bool unitarize_link(const M3 &V, M3 &W, const ReunitParams rp) {
  double det_check = 0;
  if (rp.allow_svd) {   // |det V|^2, :849-880
    const double2 *e = V.e;
    auto cm = [](double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); };
    const double2 p1 = cm(e[4], e[8]), q1 = cm(e[5], e[7]);
    const double2 p2 = cm(e[3], e[8]), q2 = cm(e[5], e[6]);
    const double2 p3 = cm(e[3], e[7]), q3 = cm(e[4], e[6]);
    const double2 a1 = make_double2(p1.x - q1.x, p1.y - q1.y), a2 = make_double2(p2.x - q2.x, p2.y - q2.y),
                  a3 = make_double2(p3.x - q3.x, p3.y - q3.y);
    const double2 d1 = cm(e[0], a1), d2 = cm(e[1], a2), d3 = cm(e[2], a3);
    const double dr = d1.x - d2.x + d3.x, di = d1.y - d2.y + d3.y;
    det_check = dr * dr + di * di;
  }
  const M3 Q = m3_an(V, V);
  const M3 Q2 = m3_nn(Q, Q);
  double q3d = 0;   // Re tr Q^3: only the diagonal of Q^3 is needed, :962-972
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) q3d += Q2.e[3 * i + k].x * Q.e[3 * k + i].x - Q2.e[3 * i + k].y * Q.e[3 * k + i].y;
  const double c0 = Q.e[0].x + Q.e[4].x + Q.e[8].x;
  const double c1 = (Q2.e[0].x + Q2.e[4].x + Q2.e[8].x) / 2;
  const double c2 = q3d / 3;
  double S = c1 / 3 - c0 * (c0 / 18);
  double g0, g1, g2;
  if (fabs(S) < 1.0e-14) {   // U3_UNIT_ANALYTIC_EPS
    g0 = g1 = g2 = c0 / 3;
  } else {
    const double R = c2 / 2 - c0 * (c1 / 3) + c0 * c0 * (c0 / 27);
    S = sqrt(S);
    const double RoS = R / (S * S * S);
    double theta;
    if (!(fabs(RoS) < 1.0)) theta = (R > 0) ? 0.0 : 3.14159265358979323846264338328;
    else theta = acos(RoS);
    const double theta3 = theta / 3, pi23 = 6.28318530717958647692528676656 / 3;
    g0 = c0 / 3 + 2 * S * cos(theta3);
    g1 = c0 / 3 + 2 * S * cos(theta3 + pi23);
    g2 = c0 / 3 + 2 * S * cos(theta3 + 2 * pi23);
  }
  if (rp.allow_svd) {   // :1039-1053
    bool svd = false;
    if (det_check != 0 && fabs(det_check - g0 * g1 * g2) / fabs(det_check) > rp.svd_rel) svd = true;
    if (det_check < rp.svd_abs) svd = true;
    if (svd) {
      W = polar_by_svd(V);
      return true;
    }
  }
  const double g0sq = sqrt(g0), g1sq = sqrt(g1), g2sq = sqrt(g2);
  double us = g1sq + g2sq, ws = g1sq * g2sq;
  const double vs = g0sq * us + ws;
  us += g0sq;
  ws *= g0sq;
  const double denom = ws * (us * vs - ws);
  const double f0 = (us * vs * vs - ws * (us * us + vs)) / denom;
  const double f1 = (2 * us * vs - ws - us * us * us) / denom;
  const double f2 = us / denom;
  M3 S2;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    S2.e[k].x = f1 * Q.e[k].x + f2 * Q2.e[k].x;
    S2.e[k].y = f1 * Q.e[k].y + f2 * Q2.e[k].y;
  }
  S2.e[0].x += f0;
  S2.e[4].x += f0;
  S2.e[8].x += f0;
  W = m3_nn(V, S2);
  return false;
}

// W = U(3) projection of the four links of every site; *nsvd counts the SVD branch
__global__ void __launch_bounds__(kBlock)
unitarize_kernel(double2 *W, const double2 *V, size_t fstride, int nsites, ReunitParams rp, unsigned long long *nsvd) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k >= 4 * nsites) return;
  const int mu = k / nsites, f = k - mu * nsites;
  const M3 v = m3_load(V + (size_t)mu * 9 * fstride, fstride, f);
  M3 w;
  if (unitarize_link(v, w, rp)) atomicAdd(nsvd, 1ull);
  m3_store(W + (size_t)mu * 9 * fstride, fstride, f, w);
}

// thin links with KS phases and the antiperiodic time boundary folded in, Haar-random SU(3)
// (synth.cuh thin_link), full-lattice layout: benchmark input for the link construction
__global__ void __launch_bounds__(kBlock)
synth_thin_kernel(double2 *links, const Geom g, size_t fstride, int nsites, uint64_t seed) {
  const int f = blockIdx.x * kBlock + threadIdx.x;
  if (f >= nsites) return;
  const int par = f >= g.Vh ? 1 : 0;
  const Coord c = site_coord(g, f - par * g.Vh, par);
  const int x[4] = {c.x, c.y, c.z, c.t};
  const double eta[4] = {(x[3] & 1) ? -1.0 : 1.0, ((x[3] + x[0]) & 1) ? -1.0 : 1.0,
                         ((x[3] + x[0] + x[1]) & 1) ? -1.0 : 1.0, 1.0};
  for (int mu = 0; mu < 4; mu++) {
    Cplx U[9];
    thin_link(seed, x, g.G, mu, U);
    double s = eta[mu];
    if (mu == 3 && x[3] == g.G[3] - 1) s = -s;
    for (int e = 0; e < 9; e++) links[(size_t)(mu * 9 + e) * fstride + f] = make_double2(s * U[e].x, s * U[e].y);
  }
}

}  // namespace b200ks
