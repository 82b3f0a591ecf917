// force.cuh -- HISQ fermion force as reverse-mode differentiation of
// the link construction of links.cuh.
//
// What the reference computes (generic_ks/fermion_force_hisq_multi.c:1183-1476; GPU seam
// qudaHisqForce, :2169-2290): for S = sum_j res_j |D_oe[U] X_j|^2, with the HISQ chain
// U -> V (fat7) -> W (U(3) projection) -> (fat, long) inside D, the momentum increment
//     A_mu(x) = -eps * TA( U_mu(x) G_U,mu(x)^+ ),      dS = Re tr(G_U^+ dU),
// TA = traceless anti-Hermitian part.  The reference walks sorted path tables; here the same
// derivative is taken by running the forward chain's staple passes backwards ("G_M" below is the
// gradient matrix of S with respect to M, dS = Re tr(G_M^+ dM)):
//   1. outer products       G_fat,mu(x) = +-c1_j Z_j(x) Z_j(x+mu)^+ , G_lng,mu(x) = +-c3_j Z_j(x) Z_j(x+3mu)^+
//                           (Z = X on even sites, D X on odd sites; + on odd x)      cf. :2009-2154
//   2. level-2 smearing and the Naik product backwards: G_fat, G_lng -> G_W          cf. :1638-1874
//   3. U(3) projection backwards: G_W -> G_V, exact derivative of V (V^+V)^-1/2 from the
//      eigen-decomposition of V^+V (Daleckii-Krein); with the reference's force filter
//      (HISQ_FORCE_FILTER, su3_mat_op.c:1680-1734) a link whose smallest eigenvalue of V^+V is
//      below the filter gets the derivative of V (V^+V + filter)^-1/2 instead             cf. :1877-2006
//   4. level-1 (fat7) smearing backwards: G_V -> G_U                                  cf. :1638-1874
//   5. projection onto the momenta                                                    cf. :1433-1470
// A forward staple pass  S(L)(x) = U_nu(x) L(x+nu) U_nu(x+mu)^+ + U_nu(x-nu)^+ L(x-nu) U_nu(x-nu+mu)
// (generic/general_staple.c:41-123) has six gradient contributions; written as gathers at the
// destination site z (no atomics), with H the gradient w.r.t. S(L):
//   G_L(z)    += U_nu(z-nu)^+ H(z-nu) U_nu(z-nu+mu) + U_nu(z) H(z+nu) U_nu(z+mu)^+
//   G_Unu(z)  += H(z) U_nu(z+mu) L(z+nu)^+ + H(z-mu)^+ U_nu(z-mu) L(z-mu+nu)
//              + L(z-mu)^+ U_nu(z-mu) H(z-mu+nu) + L(z) U_nu(z+mu) H(z+nu)^+
//
// The Lepage term is differentiated as the path the reference's force table holds (+nu+nu+mu-nu-nu with
// the one-link coefficient as given, imp_actions/hisq/hisq_u3_action.h), i.e. the upper staple of the
// upper 3-staple plus the lower of the lower, not as the fattening computes it (the full staple of the
// full staple, whose two back-tracking terms the "one_link - 6 lepage" coefficient cancels for
// unitary links).  The two agree in every direction tangent to the group, so the force is the same;
// their radial parts differ, and a filtered link of step 3 sees the radial part of G_W.
//
// Every per-site routine is __host__ __device__: tests/host/force_host.cu runs the same bodies in
// plain host loops against the CPU oracle (oracle/ks_force_oracle.c), so the arithmetic and the
// neighbour bookkeeping are checked without a GPU; the kernels only add the launch.
// Layout as links.cuh: matrix fields of 9 double2 planes, plane stride fs, site f = parity*Vh + cb.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>

namespace b200ks {
namespace force {

#define B200KS_HD __host__ __device__ inline

struct FGeom {   // periodic single-GPU lattice; sites in MILC order (even block, then odd block)
  int L[4];
  int Vh;
};

struct Mat { double2 e[9]; };

B200KS_HD Mat ld(const double2 *p, size_t fs, int f) {
  Mat m;
  for (int k = 0; k < 9; k++) m.e[k] = p[(size_t)k * fs + f];
  return m;
}
B200KS_HD void st(double2 *p, size_t fs, int f, const Mat &m) {
  for (int k = 0; k < 9; k++) p[(size_t)k * fs + f] = m.e[k];
}
B200KS_HD void acc(double2 *p, size_t fs, int f, double s, const Mat &m) {   // p(f) += s m
  for (int k = 0; k < 9; k++) {
    double2 v = p[(size_t)k * fs + f];
    v.x += s * m.e[k].x;
    v.y += s * m.e[k].y;
    p[(size_t)k * fs + f] = v;
  }
}
B200KS_HD Mat nn(const Mat &a, const Mat &b) {
  Mat c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double re = 0, im = 0;
      for (int k = 0; k < 3; k++) {   // one product per statement: each contracts to a single DFMA on the device
        const double2 x = a.e[3 * i + k], y = b.e[3 * k + j];
        re += x.x * y.x;
        re -= x.y * y.y;
        im += x.x * y.y;
        im += x.y * y.x;
      }
      c.e[3 * i + j] = make_double2(re, im);
    }
  return c;
}
B200KS_HD Mat dag(const Mat &a) {
  Mat c;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c.e[3 * i + j] = make_double2(a.e[3 * j + i].x, -a.e[3 * j + i].y);
  return c;
}
B200KS_HD Mat na(const Mat &a, const Mat &b) { return nn(a, dag(b)); }
B200KS_HD Mat an(const Mat &a, const Mat &b) { return nn(dag(a), b); }
B200KS_HD void add(Mat &a, const Mat &b) {
  for (int k = 0; k < 9; k++) {
    a.e[k].x += b.e[k].x;
    a.e[k].y += b.e[k].y;
  }
}
B200KS_HD Mat zero() {
  Mat m;
  for (int k = 0; k < 9; k++) m.e[k] = make_double2(0.0, 0.0);
  return m;
}

// site f = parity*Vh + cb, cb = lex/2.  A site's coordinates cost three integer divisions by run-time
// extents (~100 instructions); the first version paid them again for every one of the 6-10 neighbours of
// a staple pass -- more instructions than the six 3x3 complex products the pass is about.  They are
// computed ONCE per site (fsite) and a hop only moves one coordinate (fhop: adds and selects, no indexed
// array, which would live in local memory).
struct FSite {
  int par, lex, c0, c1, c2, c3;
};
B200KS_HD FSite fsite(const FGeom &g, int f) {
  FSite s;
  s.par = f >= g.Vh ? 1 : 0;
  const int cb = f - s.par * g.Vh;
  const int Lxh = g.L[0] / 2;
  int r = cb;
  const int xh = r % Lxh;
  r /= Lxh;
  s.c1 = r % g.L[1];
  r /= g.L[1];
  s.c2 = r % g.L[2];
  s.c3 = r / g.L[2];
  s.c0 = 2 * xh + ((s.c1 + s.c2 + s.c3 + s.par) & 1);
  s.lex = s.c0 + g.L[0] * (s.c1 + g.L[1] * (s.c2 + g.L[2] * s.c3));
  return s;
}
// the site at +-1 in direction mu (periodic)
B200KS_HD FSite fhop(const FGeom &g, const FSite &s, int mu, int sign) {
  const int cm = mu == 0 ? s.c0 : mu == 1 ? s.c1 : mu == 2 ? s.c2 : s.c3;
  const int Lm = mu == 0 ? g.L[0] : mu == 1 ? g.L[1] : mu == 2 ? g.L[2] : g.L[3];
  const int sm = mu == 0 ? 1 : mu == 1 ? g.L[0] : mu == 2 ? g.L[0] * g.L[1] : g.L[0] * g.L[1] * g.L[2];
  int cn = cm + sign;
  if (cn >= Lm) cn -= Lm;
  if (cn < 0) cn += Lm;
  FSite t = s;
  t.par = s.par ^ 1;
  t.lex = s.lex + (cn - cm) * sm;
  t.c0 = mu == 0 ? cn : s.c0;
  t.c1 = mu == 1 ? cn : s.c1;
  t.c2 = mu == 2 ? cn : s.c2;
  t.c3 = mu == 3 ? cn : s.c3;
  return t;
}
B200KS_HD int findex(const FGeom &g, const FSite &s) { return s.par * g.Vh + (s.lex >> 1); }
B200KS_HD int nbr(const FGeom &g, int f, int mu, int sign) { return findex(g, fhop(g, fsite(g, f), mu, sign)); }

// ---- site functors (operator()(f) for every site f) ------------------------------------------------
// 1. outer products of one term: z = colour vectors in MILC host order, 6 reals per site
struct OprodSite {
  FGeom g;
  double2 *gfat, *glng;
  size_t fs;
  const double *z;
  double c1, c3;
  B200KS_HD void operator()(int f) const {
    const double sgn = f >= g.Vh ? 1.0 : -1.0;
    const FSite s0 = fsite(g, f);
    for (int mu = 0; mu < 4; mu++) {
      const FSite s1 = fhop(g, s0, mu, 1);
      const int f1 = findex(g, s1), f3 = findex(g, fhop(g, fhop(g, s1, mu, 1), mu, 1));
      Mat o1, o3;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          const double zr = z[6 * (size_t)f + 2 * a], zi = z[6 * (size_t)f + 2 * a + 1];
          const double pr = z[6 * (size_t)f1 + 2 * b], pi = z[6 * (size_t)f1 + 2 * b + 1];
          const double qr = z[6 * (size_t)f3 + 2 * b], qi = z[6 * (size_t)f3 + 2 * b + 1];
          o1.e[3 * a + b] = make_double2(zr * pr + zi * pi, zi * pr - zr * pi);
          o3.e[3 * a + b] = make_double2(zr * qr + zi * qi, zi * qr - zr * qi);
        }
      acc(gfat + (size_t)mu * 9 * fs, fs, f, sgn * c1, o1);
      acc(glng + (size_t)mu * 9 * fs, fs, f, sgn * c3, o3);
    }
  }
};

// 2. forward staple (no fat-link accumulation): out = S(link); part 1 = the upper staple only
//    (U_nu(x) L(x+nu) U_nu(x+mu)^+), 2 = the lower one only, 3 = both
struct StapleFwdSite {
  FGeom g;
  double2 *out;
  const double2 *link, *Unu;   // matrix fields: the mu "link" and the gauge links of direction nu
  size_t fs;
  int mu, nu;
  int part = 3;
  B200KS_HD void operator()(int f) const {
    Mat r = zero();
    const FSite s0 = fsite(g, f);
    if (part & 1) {
      const int fpn = findex(g, fhop(g, s0, nu, 1)), fpm = findex(g, fhop(g, s0, mu, 1));
      r = nn(ld(Unu, fs, f), na(ld(link, fs, fpn), ld(Unu, fs, fpm)));
    }
    if (part & 2) {
      const FSite smn = fhop(g, s0, nu, -1);
      const int fmn = findex(g, smn), fmnpm = findex(g, fhop(g, smn, mu, 1));
      add(r, nn(an(ld(Unu, fs, fmn), ld(link, fs, fmn)), ld(Unu, fs, fmnpm)));
    }
    st(out, fs, f, r);
  }
};

// 3. backward staple: H = hs * Hfield is the gradient w.r.t. S(link); adds to glink and gUnu at z.
//    part as in StapleFwdSite: the contributions of the upper staple (1), the lower one (2) or both (3)
#ifndef B200KS_FORCE_BWD_MINB
#define B200KS_FORCE_BWD_MINB 2   // CTAs per SM the fused backward staple body is compiled for (A/B: profiles/variants)
#endif
struct StapleBwdSite {
  static constexpr int kMinBlocks = B200KS_FORCE_BWD_MINB;
  FGeom g;
  const double2 *H;
  double hs;
  const double2 *link, *Unu;
  double2 *glink, *gUnu;
  size_t fs;
  int mu, nu;
  int part = 3;
  B200KS_HD void operator()(int z) const {
    const FSite s0 = fsite(g, z), smn = fhop(g, s0, nu, -1), smm = fhop(g, s0, mu, -1);
    const int zpn = findex(g, fhop(g, s0, nu, 1)), zmn = findex(g, smn), zpm = findex(g, fhop(g, s0, mu, 1)), zmm = findex(g, smm);
    const int zmnpm = findex(g, fhop(g, smn, mu, 1)), zmmpn = findex(g, fhop(g, smm, nu, 1));
    const Mat Uzpm = ld(Unu, fs, zpm), Uzmm = ld(Unu, fs, zmm);
    Mat gl = zero(), gu = zero();
    if (part & 1) {   // upper staple A B C^+ at x: G_B(x+nu), G_A(x), G_C(x+mu)
      gl = nn(an(ld(Unu, fs, zmn), ld(H, fs, zmn)), ld(Unu, fs, zmnpm));
      gu = nn(ld(H, fs, z), na(Uzpm, ld(link, fs, zpn)));
      add(gu, nn(an(ld(H, fs, zmm), Uzmm), ld(link, fs, zmmpn)));
    }
    if (part & 2) {   // lower staple D^+ E F at x, y = x - nu: G_E(y), G_F(y+mu), G_D(y)
      const Mat Hzpn = ld(H, fs, zpn);
      add(gl, na(nn(ld(Unu, fs, z), Hzpn), Uzpm));
      add(gu, nn(an(ld(link, fs, zmm), Uzmm), ld(H, fs, zmmpn)));
      add(gu, na(nn(ld(link, fs, z), Uzpm), Hzpn));
    }
    acc(glink, fs, z, hs, gl);
    acc(gUnu, fs, z, hs, gu);
  }
};

// The same pass as up to four kernels (gradient of the link field / of the nu links, upper / lower staple, the
// staple chosen at compile time): a few more matrix loads per site in total, but each piece keeps 3 to 6
// matrices live instead of the fused body's 14 (254 registers: 8 warps per SM) and is compiled for 128 registers
// (kMinBlocks, used by the launch in fermion_force.cu: 16 warps per SM).  Selected at run time (ForceBufs::split) for A/B measurements.
template <int kPart>
struct StapleBwdLinkSite {
  static constexpr int kMinBlocks = 4;
  FGeom g;
  const double2 *H;
  double hs;
  const double2 *Unu;
  double2 *glink;
  size_t fs;
  int mu, nu;
  B200KS_HD void operator()(int z) const {
    if (kPart == 1) {
      const FSite smn = fhop(g, fsite(g, z), nu, -1);
      const int zmn = findex(g, smn), zmnpm = findex(g, fhop(g, smn, mu, 1));
      acc(glink, fs, z, hs, nn(an(ld(Unu, fs, zmn), ld(H, fs, zmn)), ld(Unu, fs, zmnpm)));
    } else {
      const FSite s0 = fsite(g, z);
      const int zpn = findex(g, fhop(g, s0, nu, 1)), zpm = findex(g, fhop(g, s0, mu, 1));
      acc(glink, fs, z, hs, na(nn(ld(Unu, fs, z), ld(H, fs, zpn)), ld(Unu, fs, zpm)));
    }
  }
};
template <int kPart>
struct StapleBwdUSite {
  static constexpr int kMinBlocks = 4;
  FGeom g;
  const double2 *H;
  double hs;
  const double2 *link, *Unu;
  double2 *gUnu;
  size_t fs;
  int mu, nu;
  B200KS_HD void operator()(int z) const {
    const FSite s0 = fsite(g, z), smm = fhop(g, s0, mu, -1);
    const int zpn = findex(g, fhop(g, s0, nu, 1)), zpm = findex(g, fhop(g, s0, mu, 1)), zmm = findex(g, smm), zmmpn = findex(g, fhop(g, smm, nu, 1));
    Mat gu;
    if (kPart == 1) {
      gu = nn(ld(H, fs, z), na(ld(Unu, fs, zpm), ld(link, fs, zpn)));
      add(gu, nn(an(ld(H, fs, zmm), ld(Unu, fs, zmm)), ld(link, fs, zmmpn)));
    } else {
      gu = nn(an(ld(link, fs, zmm), ld(Unu, fs, zmm)), ld(H, fs, zmmpn));
      add(gu, na(nn(ld(link, fs, z), ld(Unu, fs, zpm)), ld(H, fs, zpn)));
    }
    acc(gUnu, fs, z, hs, gu);
  }
};

// The full pass as two ROLES of one kernel (fermion_force.cu force_pair_kernel): two threads per site, each with three
// of the six contributions and ONE of the two outputs, so that a role keeps about half of the fused body's matrices
// live and compiles for 128 registers without spilling (16 warps per SM instead of 8; the fused body at 128 registers
// spills 2 KB per thread; the roles spill 0.2 KB).  Unlike the four-kernel split the two roles of a site run in the same CTA at the same
// time: the matrices both of them load (U_nu(z+mu), H(z+nu)) and the neighbours' reloads meet in L1/L2, not in HBM.
//   role_link: both contributions to the gradient of the link field -> glink(z); returns the sixth contribution
//              (L(z) U_nu(z+mu) H(z+nu)^+, whose factors it holds anyway), handed over through shared memory
//   role_u:    the other three contributions to the gradient of the nu links; finish_u adds the sixth, -> gUnu(z)
// Same products, same order of summation as StapleBwdSite with part = 3: the results are identical.
template <int kMB>   // CTAs per SM the kernel is compiled for: 4 = 128 registers (16 warps per SM), 3 = 168 (12 warps)
struct StapleBwdPairSite {
  static constexpr int kMinBlocks = kMB;
  FGeom g;
  const double2 *H;
  double hs;
  const double2 *link, *Unu;
  double2 *glink, *gUnu;
  size_t fs;
  int mu, nu;
  // hand(t4) is called as soon as the sixth contribution is formed (nothing of it stays live afterwards)
  template <class Hand>
  B200KS_HD void role_link(int z, const Hand &hand) const {
    const FSite s0 = fsite(g, z), smn = fhop(g, s0, nu, -1);
    const int zpn = findex(g, fhop(g, s0, nu, 1)), zpm = findex(g, fhop(g, s0, mu, 1));
    const int zmn = findex(g, smn), zmnpm = findex(g, fhop(g, smn, mu, 1));
    hand(na(nn(ld(link, fs, z), ld(Unu, fs, zpm)), ld(H, fs, zpn)));
    Mat gl = na(nn(ld(Unu, fs, z), ld(H, fs, zpn)), ld(Unu, fs, zpm));
    const Mat up = nn(an(ld(Unu, fs, zmn), ld(H, fs, zmn)), ld(Unu, fs, zmnpm));
    for (int k = 0; k < 9; k++) {   // upper + lower, the fused body's order
      gl.e[k].x = up.e[k].x + gl.e[k].x;
      gl.e[k].y = up.e[k].y + gl.e[k].y;
    }
    acc(glink, fs, z, hs, gl);
  }
  // sum(term) is called with the three contributions one after the other (the device sums them in shared memory: a
  // 3x3 accumulator held in registers across the terms is what does not fit into 128 of them)
  template <class Sum>
  B200KS_HD void role_u(int z, const Sum &sum) const {
    const FSite s0 = fsite(g, z), smm = fhop(g, s0, mu, -1);
    const int zpn = findex(g, fhop(g, s0, nu, 1)), zpm = findex(g, fhop(g, s0, mu, 1));
    const int zmm = findex(g, smm), zmmpn = findex(g, fhop(g, smm, nu, 1));
    sum(nn(ld(H, fs, z), na(ld(Unu, fs, zpm), ld(link, fs, zpn))));
    sum(nn(an(ld(H, fs, zmm), ld(Unu, fs, zmm)), ld(link, fs, zmmpn)));
    sum(nn(an(ld(link, fs, zmm), ld(Unu, fs, zmm)), ld(H, fs, zmmpn)));
  }
  B200KS_HD void finish_u(int z, Mat gu, const Mat &t4) const {
    add(gu, t4);
    acc(gUnu, fs, z, hs, gu);
  }
  struct Keep {   // (host executor: hand-over and sum are local variables)
    Mat *m;
    B200KS_HD void operator()(const Mat &t) const { *m = t; }
  };
  struct Add {
    Mat *m;
    B200KS_HD void operator()(const Mat &t) const { add(*m, t); }
  };
  B200KS_HD void operator()(int z) const {   // (host executor; the device runs the roles on two threads)
    Mat t4, gu = zero();
    role_link(z, Keep{&t4});
    role_u(z, Add{&gu});
    finish_u(z, gu, t4);
  }
};

// out(f) += s * in(f) for nplanes planes (one-link term backwards, scaled copies)
struct AxpySite {
  double2 *out;
  const double2 *in;
  double s;
  size_t fs;
  int nplanes;
  B200KS_HD void operator()(int f) const {
    for (int k = 0; k < nplanes; k++) {
      double2 v = out[(size_t)k * fs + f];
      const double2 w = in[(size_t)k * fs + f];
      v.x += s * w.x;
      v.y += s * w.y;
      out[(size_t)k * fs + f] = v;
    }
  }
};
struct ZeroSite {
  double2 *out;
  size_t fs;
  int nplanes;
  B200KS_HD void operator()(int f) const {
    for (int k = 0; k < nplanes; k++) out[(size_t)k * fs + f] = make_double2(0.0, 0.0);
  }
};

// 4. Naik product backwards (lng = s * W W W): gW(z) += s [ G(z) (W1 W2)^+ + W(-1)^+ G(-1) W1^+ + (W(-2) W(-1))^+ G(-2) ]
struct NaikBwdSite {
  FGeom g;
  const double2 *glng, *W;   // four directions, 36 planes each
  double2 *gW;
  double s;
  size_t fs;
  B200KS_HD void operator()(int z) const {
    const FSite s0 = fsite(g, z);
    for (int mu = 0; mu < 4; mu++) {
      const double2 *Wm = W + (size_t)mu * 9 * fs, *Gm = glng + (size_t)mu * 9 * fs;
      const FSite sp1 = fhop(g, s0, mu, 1), sm1 = fhop(g, s0, mu, -1);
      const int p1 = findex(g, sp1), p2 = findex(g, fhop(g, sp1, mu, 1)), m1 = findex(g, sm1), m2 = findex(g, fhop(g, sm1, mu, -1));
      const Mat W1 = ld(Wm, fs, p1), Wm1 = ld(Wm, fs, m1);
      Mat r = na(ld(Gm, fs, z), nn(W1, ld(Wm, fs, p2)));
      add(r, na(an(Wm1, ld(Gm, fs, m1)), W1));
      add(r, an(nn(ld(Wm, fs, m2), Wm1), ld(Gm, fs, m2)));
      acc(gW + (size_t)mu * 9 * fs, fs, z, s, r);
    }
  }
};

// Hermitian 3x3 eigen-decomposition by cyclic Jacobi rotations: Q = E diag(g) E^+
B200KS_HD void herm_eig(const Mat &Q, double (&g)[3], Mat &E) {
  double2 a[3][3], v[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      a[i][j] = Q.e[3 * i + j];
      v[i][j] = make_double2(i == j ? 1.0 : 0.0, 0.0);
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) off += a[p][q].x * a[p][q].x + a[p][q].y * a[p][q].y;
    if (off < 1e-60) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double ar = a[p][q].x, ai = a[p][q].y, ab = sqrt(ar * ar + ai * ai);
        if (ab < 1e-300) continue;
        const double er = ar / ab, ei = ai / ab;
        const double theta = (a[q][q].x - a[p][p].x) / (2.0 * ab);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++)
          for (int w = 0; w < 2; w++) {   // columns p, q of a and of v
            double2 &P = w ? v[i][p] : a[i][p];
            double2 &R = w ? v[i][q] : a[i][q];
            const double pr = P.x, pi = P.y, qr = R.x, qi = R.y;
            const double cqr = er * qr + ei * qi, cqi = er * qi - ei * qr;
            const double epr = er * pr - ei * pi, epi = er * pi + ei * pr;
            P = make_double2(c * pr - s * cqr, c * pi - s * cqi);
            R = make_double2(s * epr + c * qr, s * epi + c * qi);
          }
        for (int j = 0; j < 3; j++) {     // rows p, q of a
          const double pr = a[p][j].x, pi = a[p][j].y, qr = a[q][j].x, qi = a[q][j].y;
          const double eqr = er * qr - ei * qi, eqi = er * qi + ei * qr;
          const double cpr = er * pr + ei * pi, cpi = er * pi - ei * pr;
          a[p][j] = make_double2(c * pr - s * eqr, c * pi - s * eqi);
          a[q][j] = make_double2(s * cpr + c * qr, s * cpi + c * qi);
        }
      }
  }
  for (int i = 0; i < 3; i++) g[i] = a[i][i].x;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) E.e[3 * i + j] = v[i][j];
}

// 5. U(3) projection backwards, one link:  G_V = G_W Q^-1/2 + V (G_Q + G_Q^+),
//    G_Q = E [phi .* (E^+ (V^+ G_W) E)] E^+,  phi_ij = (g_i^-1/2 - g_j^-1/2)/(g_i - g_j)
//    filter > 0: when the smallest g_i is below it, all three are shifted by it (see the header)
B200KS_HD Mat unit_bwd(const Mat &V, const Mat &GW, double filter) {
  const Mat Q = an(V, V);
  double g[3];
  Mat E;
  herm_eig(Q, g, E);
  if (filter > 0.0 && fmin(g[0], fmin(g[1], g[2])) < filter)
    for (int i = 0; i < 3; i++) g[i] += filter;
  const Mat Ed = dag(E);
  Mat Rt = nn(Ed, nn(an(V, GW), E));
  Mat S = zero();
  for (int i = 0; i < 3; i++) {
    const double ri = sqrt(g[i]);
    S.e[4 * i].x = 1.0 / ri;
    for (int j = 0; j < 3; j++) {
      const double rj = sqrt(g[j]);
      const double phi = -1.0 / (ri * rj * (ri + rj));   // = (1/ri - 1/rj)/(g_i - g_j), also at g_i = g_j
      Rt.e[3 * i + j].x *= phi;
      Rt.e[3 * i + j].y *= phi;
    }
  }
  Mat Gq = nn(E, nn(Rt, Ed));
  add(Gq, dag(Gq));
  Mat r = nn(GW, nn(E, nn(S, Ed)));
  add(r, nn(V, Gq));
  return r;
}
struct UnitBwdSite {   // index k over 4*nsites links: mu = k / nsites
  const double2 *V;
  double2 *gW;         // in: G_W, out: G_V (in place)
  size_t fs;
  int nsites;
  double filter;
  B200KS_HD void operator()(int k) const {
    const int mu = k / nsites, f = k - mu * nsites;
    const size_t o = (size_t)mu * 9 * fs;
    st(gW + o, fs, f, unit_bwd(ld(V + o, fs, f), ld(gW + o, fs, f), filter));
  }
};

// 6. momentum increment A = -eps TA(U G_U^+) as MILC's anti_hermitmat (include/su3.h):
//    {m01.re, m01.im, m02.re, m02.im, m12.re, m12.im, m00im, m11im, m22im, space}
template <typename TH>
struct MomSite {
  const double2 *U, *gU;
  TH *mom;             // [site][dir][10]
  double eps;
  size_t fs;
  int nsites;
  B200KS_HD void operator()(int k) const {
    const int mu = k / nsites, f = k - mu * nsites;
    const size_t o = (size_t)mu * 9 * fs;
    const Mat M = na(ld(U + o, fs, f), ld(gU + o, fs, f));
    const double tr = (M.e[0].y + M.e[4].y + M.e[8].y) / 3.0;
    TH *m = mom + 10 * ((size_t)4 * f + mu);
    m[0] = (TH)(-eps * 0.5 * (M.e[1].x - M.e[3].x));
    m[1] = (TH)(-eps * 0.5 * (M.e[1].y + M.e[3].y));
    m[2] = (TH)(-eps * 0.5 * (M.e[2].x - M.e[6].x));
    m[3] = (TH)(-eps * 0.5 * (M.e[2].y + M.e[6].y));
    m[4] = (TH)(-eps * 0.5 * (M.e[5].x - M.e[7].x));
    m[5] = (TH)(-eps * 0.5 * (M.e[5].y + M.e[7].y));
    m[6] = (TH)(-eps * (M.e[0].y - tr));
    m[7] = (TH)(-eps * (M.e[4].y - tr));
    m[8] = (TH)(-eps * (M.e[8].y - tr));
    m[9] = (TH)0;
  }
};

// ---- the chain, written once for any executor X with  template <class F> void X::run(int n, const F &f) ----
struct ForceBufs {
  FGeom g;
  size_t fs;          // plane stride of every matrix field
  int nsites;
  double2 *U, *V, *W;             // 36 planes each (inputs)
  double2 *gfat, *glng, *gW, *gU; // 36 planes each
  double2 *st3, *st5, *g3, *g5;   // 9 planes each
  bool split = false;             // backward staple passes as up to four small kernels (StapleBwdLinkSite / StapleBwdUSite)
  int pair = 0;                   // full backward staple passes as two roles of one kernel (StapleBwdPairSite<pair>: 3 or 4)
};

template <class X>
void staple_bwd(X &x, const ForceBufs &b, const double2 *H, double hs, const double2 *link, const double2 *Unu, double2 *glink,
                double2 *gUnu, int mu, int nu, int part = 3) {
  if (b.split) {
    if (part & 1) {
      x.run(b.nsites, StapleBwdLinkSite<1>{b.g, H, hs, Unu, glink, b.fs, mu, nu});
      x.run(b.nsites, StapleBwdUSite<1>{b.g, H, hs, link, Unu, gUnu, b.fs, mu, nu});
    }
    if (part & 2) {
      x.run(b.nsites, StapleBwdLinkSite<2>{b.g, H, hs, Unu, glink, b.fs, mu, nu});
      x.run(b.nsites, StapleBwdUSite<2>{b.g, H, hs, link, Unu, gUnu, b.fs, mu, nu});
    }
  } else if (b.pair == 4 && part == 3) {
    x.run(b.nsites, StapleBwdPairSite<4>{b.g, H, hs, link, Unu, glink, gUnu, b.fs, mu, nu});
  } else if (b.pair == 3 && part == 3) {
    x.run(b.nsites, StapleBwdPairSite<3>{b.g, H, hs, link, Unu, glink, gUnu, b.fs, mu, nu});
  } else {
    x.run(b.nsites, StapleBwdSite{b.g, H, hs, link, Unu, glink, gUnu, b.fs, mu, nu, part});
  }
}

// reverse of one smearing level (links.cuh smear_dev / load_fatlinks_cpu): gfat -> adds to glinks
template <class X>
void smear_bwd(X &x, const ForceBufs &b, const double *coeffs, const double2 *links, const double2 *gfat, double2 *glinks) {
  const double one_link = coeffs[0], three = coeffs[2], five = coeffs[3], seven = coeffs[4], lepage = coeffs[5];
  const size_t fs = b.fs, m1 = 9 * fs;
  const int n = b.nsites;
  x.run(n, AxpySite{glinks, gfat, one_link, fs, 36});
  if (three == 0.0 && lepage == 0.0 && five == 0.0) return;
  for (int dir = 0; dir < 4; dir++) {
    const double2 *Gd = gfat + dir * m1;
    for (int nu = 0; nu < 4; nu++) {
      if (nu == dir) continue;
      const double2 *Unu = links + nu * m1;
      x.run(n, StapleFwdSite{b.g, b.st3, links + dir * m1, Unu, fs, dir, nu});
      x.run(n, ZeroSite{b.g3, fs, 9});
      x.run(n, AxpySite{b.g3, Gd, three, fs, 9});
      if (lepage != 0.0)   // straight double staples: upper of upper, lower of lower (st5 / g5 are free here)
        for (int part = 1; part <= 2; part++) {
          x.run(n, StapleFwdSite{b.g, b.st5, links + dir * m1, Unu, fs, dir, nu, part});
          x.run(n, ZeroSite{b.g5, fs, 9});
          staple_bwd(x, b, Gd, lepage, b.st5, Unu, b.g5, glinks + nu * m1, dir, nu, part);
          staple_bwd(x, b, b.g5, 1.0, links + dir * m1, Unu, glinks + dir * m1, glinks + nu * m1, dir, nu, part);
        }
      for (int rho = 0; rho < 4; rho++) {
        if (rho == dir || rho == nu) continue;
        const double2 *Urho = links + rho * m1;
        x.run(n, StapleFwdSite{b.g, b.st5, b.st3, Urho, fs, dir, rho});
        x.run(n, ZeroSite{b.g5, fs, 9});
        x.run(n, AxpySite{b.g5, Gd, five, fs, 9});
        for (int sig = 0; sig < 4; sig++) {
          if (sig == dir || sig == nu || sig == rho) continue;
          staple_bwd(x, b, Gd, seven, b.st5, links + sig * m1, b.g5, glinks + sig * m1, dir, sig);
        }
        staple_bwd(x, b, b.g5, 1.0, b.st3, Urho, b.g3, glinks + rho * m1, dir, rho);
      }
      staple_bwd(x, b, b.g3, 1.0, links + dir * m1, Unu, glinks + dir * m1, glinks + nu * m1, dir, nu);
    }
  }
}

// gfat / glng hold the outer products on entry; on exit gU holds G_U.  naik_in_oprod: the three-hop
// coefficients already carry the Naik coefficient (qudaHisqForce's convention), else coeffs2[1] is applied here.
// Several Naik epsilons (fermion_force_hisq_multi.c:1285-1375): every term goes through the level-2 smearing,
// and the terms solved with a Naik epsilon add  eps_k (c1' G_fat^(k) + Naik product backwards of c3' G_lng^(k))
// straight to G_W (c1', c3': the reference's one-link + Naik table).  naik_terms: on entry gW holds the sum of
// their one-hop outer products and gU the sum of their three-hop ones, weights eps_k c1' 2 res_j and
// eps_k c3' 2 res_j included (the seam's coeff[num_terms + i]); otherwise both are cleared here.
// In two phases, so that the host side can bring V and U to the device while the first one runs:
//   force_chain_w   needs W only (level-2 smearing and the Naik products backwards): G_fat, G_lng -> G_W
//   force_chain_vu  needs V, then U (projection and level-1 smearing backwards):     G_W -> G_V -> G_U
template <class X>
void force_chain_w(X &x, const ForceBufs &b, const double *coeffs2, bool naik_in_oprod, bool naik_terms = false) {
  const int n = b.nsites;
  if (naik_terms) {
    x.run(n, NaikBwdSite{b.g, b.gU, b.W, b.gW, 1.0, b.fs});
  } else {
    x.run(n, ZeroSite{b.gW, b.fs, 36});
  }
  x.run(n, ZeroSite{b.gU, b.fs, 36});
  smear_bwd(x, b, coeffs2, b.W, b.gfat, b.gW);
  x.run(n, NaikBwdSite{b.g, b.glng, b.W, b.gW, naik_in_oprod ? 1.0 : coeffs2[1], b.fs});
}
template <class X>
void force_chain_vu(X &x, const ForceBufs &b, const double *coeffs1, double filter) {
  x.run(4 * b.nsites, UnitBwdSite{b.V, b.gW, b.fs, b.nsites, filter});
  smear_bwd(x, b, coeffs1, b.U, b.gW, b.gU);
}
template <class X>
void force_chain(X &x, const ForceBufs &b, const double *coeffs1, const double *coeffs2, bool naik_in_oprod, double filter,
                 bool naik_terms = false) {
  force_chain_w(x, b, coeffs2, naik_in_oprod, naik_terms);
  force_chain_vu(x, b, coeffs1, filter);
}

}  // namespace force
}  // namespace b200ks
