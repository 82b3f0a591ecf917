// synth.cuh -- synthetic HISQ-like fields generated on the device.
//
// For the large-volume configurations (64^3x96 and up) the host cannot hold the double
// precision links, so the benchmark inputs are produced in place by
// a counter-based generator keyed on the GLOBAL site coordinates: the fields are identical
// for every decomposition of the lattice, which is what makes strong-scaling runs solve the
// same problem at every GPU count.  Structure follows SURVEY.md 8(d) config 2 (same as
// milc_qcd_b200/fields.py): thin links U Haar SU(3); fat = eta*(U + noise), long =
// eta*c3*U(x)U(x+mu)U(x+2mu), antiperiodic time boundary signs folded in
// (generic_ks/fermion_links_fn_twist_milc.c:137-141, generic_ks/rephase.c:83-115).
#pragma once
#include "common.cuh"
#include "layout.cuh"

namespace b200ks {

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// k-th standard normal pair of stream `key`
__device__ __forceinline__ double2 gauss_pair(uint64_t key, uint32_t k) {
  const uint64_t a = splitmix64(key + 2ull * k), b = splitmix64(key + 2ull * k + 1ull);
  const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0,1]
  const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
  const double r = sqrt(-2.0 * log(u1));
  double sn, cs;
  sincospi(2.0 * u2, &sn, &cs);
  return make_double2(r * cs, r * sn);
}

struct Cplx { double x, y; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ Cplx cconj(Cplx a) { return {a.x, -a.y}; }

// Haar SU(3) from two Gaussian rows (Gram-Schmidt, third row = conjugate cross product)
__device__ inline void thin_link(uint64_t seed, const int x[4], const int G[4], int mu, Cplx (&U)[9]) {
  const uint64_t lex = (uint64_t)x[0] + (uint64_t)G[0] * (x[1] + (uint64_t)G[1] * (x[2] + (uint64_t)G[2] * x[3]));
  const uint64_t key = splitmix64(seed ^ splitmix64(lex * 4ull + mu)) & ~0xFFull;
  Cplx a[3], b[3];
  double na = 0;
  for (int k = 0; k < 3; k++) {
    const double2 g = gauss_pair(key, k);
    a[k] = {g.x, g.y};
    na += g.x * g.x + g.y * g.y;
  }
  na = 1.0 / sqrt(na);
  for (int k = 0; k < 3; k++) { a[k].x *= na; a[k].y *= na; }
  Cplx dot = {0, 0};
  for (int k = 0; k < 3; k++) {
    const double2 g = gauss_pair(key, 3 + k);
    b[k] = {g.x, g.y};
    const Cplx t = cmul(cconj(a[k]), b[k]);
    dot.x += t.x; dot.y += t.y;
  }
  double nb = 0;
  for (int k = 0; k < 3; k++) {
    const Cplx t = cmul(dot, a[k]);
    b[k].x -= t.x; b[k].y -= t.y;
    nb += b[k].x * b[k].x + b[k].y * b[k].y;
  }
  nb = 1.0 / sqrt(nb);
  for (int k = 0; k < 3; k++) { b[k].x *= nb; b[k].y *= nb; }
  for (int k = 0; k < 3; k++) { U[k] = a[k]; U[3 + k] = b[k]; }
  for (int k = 0; k < 3; k++) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    const Cplx p = cmul(a[k1], b[k2]), q = cmul(a[k2], b[k1]);
    U[6 + k] = cconj(Cplx{p.x - q.x, p.y - q.y});
  }
}

__device__ inline void mat_mul(const Cplx (&A)[9], const Cplx (&B)[9], Cplx (&C)[9]) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      Cplx s = {0, 0};
      for (int k = 0; k < 3; k++) {
        const Cplx t = cmul(A[3 * r + k], B[3 * k + c]);
        s.x += t.x; s.y += t.y;
      }
      C[3 * r + c] = s;
    }
}

// local link-field site index (interior or backward-ghost tail) -> local coordinates; the
// ghost tail holds coordinates -3..-1 in the partitioned direction.
__device__ inline bool link_site_coords(const Geom &g, int i, int par, int (&x)[4]) {
  if (i < g.Vh) {
    const Coord c = site_coord(g, i, par);
    x[0] = c.x; x[1] = c.y; x[2] = c.z; x[3] = c.t;
    return true;
  }
  for (int d = 2; d < 4; d++) {
    if (!g.part[d]) continue;
    const int off = i - g.lghost[d];
    if (off < 0 || off >= 3 * g.faceh[d]) continue;
    const int slice = off / g.faceh[d], within = off - slice * g.faceh[d];
    const int S2 = g.Lxh * g.L[1];
    int xh, y, z, t;
    if (d == 3) {
      xh = within % g.Lxh;
      int r = within / g.Lxh;
      y = r % g.L[1];
      z = r / g.L[1];
      t = slice - 3;
    } else {
      t = within / S2;
      const int r2 = within - t * S2;
      xh = r2 % g.Lxh;
      y = r2 / g.Lxh;
      z = slice - 3;
    }
    x[1] = y; x[2] = z; x[3] = t;
    x[0] = 2 * xh + ((y + z + t + par) & 1);   // (z or t) = -3..-1: parity arithmetic still holds
    return true;
  }
  return false;
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
synth_links_kernel(typename Vec2<T>::type *fat, typename Vec2<T>::type *lng, Geom g, int par, uint64_t seed,
                   double fat_noise, double c3, int nsites) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= nsites) return;
  int xl[4];
  if (!link_site_coords(g, i, par, xl)) return;
  int x[4];
  for (int d = 0; d < 4; d++) x[d] = ((xl[d] + g.origin[d]) % g.G[d] + g.G[d]) % g.G[d];
  const double eta[4] = {(x[3] & 1) ? -1.0 : 1.0, ((x[3] + x[0]) & 1) ? -1.0 : 1.0,
                         ((x[3] + x[0] + x[1]) & 1) ? -1.0 : 1.0, 1.0};
  for (int mu = 0; mu < 4; mu++) {
    Cplx U0[9], U1[9], U2[9], P[9], Q[9];
    thin_link(seed, x, g.G, mu, U0);
    int y[4] = {x[0], x[1], x[2], x[3]};
    y[mu] = (x[mu] + 1) % g.G[mu];
    thin_link(seed, y, g.G, mu, U1);
    y[mu] = (x[mu] + 2) % g.G[mu];
    thin_link(seed, y, g.G, mu, U2);
    mat_mul(U0, U1, P);
    mat_mul(P, U2, Q);
    double sf = eta[mu], sl = eta[mu] * c3;
    if (mu == 3) {  // antiperiodic time boundary
      if (x[3] == g.G[3] - 1) sf = -sf;
      if (x[3] >= g.G[3] - 3) sl = -sl;
    }
    const uint64_t lex = (uint64_t)x[0] + (uint64_t)g.G[0] * (x[1] + (uint64_t)g.G[1] * (x[2] + (uint64_t)g.G[2] * x[3]));
    const uint64_t nkey = splitmix64((seed + 0x5851F42D4C957F2Dull) ^ splitmix64(lex * 4ull + mu)) & ~0xFFull;
    for (int e = 0; e < 9; e++) {
      const double2 n = gauss_pair(nkey, e);
      T2 f, l;
      f.x = (T)(sf * (U0[e].x + fat_noise * n.x));
      f.y = (T)(sf * (U0[e].y + fat_noise * n.y));
      l.x = (T)(sl * Q[e].x);
      l.y = (T)(sl * Q[e].y);
      fat[(size_t)(mu * 9 + e) * g.lstride + i] = f;
      lng[(size_t)(mu * 9 + e) * g.lstride + i] = l;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
synth_vec_kernel(typename Vec2<T>::type *v, Geom g, int par, uint64_t seed) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= g.Vh) return;
  const Coord c = site_coord(g, i, par);
  const int x[4] = {c.x + g.origin[0], c.y + g.origin[1], c.z + g.origin[2], c.t + g.origin[3]};
  const uint64_t lex = (uint64_t)x[0] + (uint64_t)g.G[0] * (x[1] + (uint64_t)g.G[1] * (x[2] + (uint64_t)g.G[2] * x[3]));
  const uint64_t key = splitmix64((seed + 0x2545F4914F6CDD1Dull) ^ splitmix64(lex)) & ~0xFFull;
  for (int q = 0; q < 3; q++) {
    const double2 n = gauss_pair(key, q);
    T2 o;
    o.x = (T)n.x;
    o.y = (T)n.y;
    v[(size_t)q * g.stride + i] = o;
  }
}

// gather the 3 highest z-slices of 9 link components of direction `dir` into a contiguous
// buffer [e][slice][t][y][xh] (link-ghost exchange in z; t slices are contiguous already)
template <typename T>
__global__ void __launch_bounds__(kBlock)
pack_zhigh_links_kernel(typename Vec2<T>::type *buf, const typename Vec2<T>::type *U, Geom g, int dir) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const int face3 = 3 * g.faceh[2];
  if (k >= 9 * face3) return;
  const int e = k / face3, r = k - e * face3;
  const int slice = r / g.faceh[2], within = r - slice * g.faceh[2];
  const int S2 = g.Lxh * g.L[1];
  const int t = within / S2, r2 = within - t * S2;
  const int z = g.L[2] - 3 + slice;
  const int idx = (t * g.L[2] + z) * S2 + r2;
  buf[k] = U[(size_t)(dir * 9 + e) * g.lstride + idx];
}

}  // namespace b200ks
