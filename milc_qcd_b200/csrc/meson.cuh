// meson.cuh -- meson tie-ups: two staggered propagators contracted into momentum-projected, time-sliced
// correlators (the step after the solves in ks_spectrum).
//
// The reference (ks_meson_cont_mom, generic_ks/ks_meson_mom.c:160-437) does, per sink spin-taste operator:
//   antiquark = O_st src1                              (spin_taste_op_fn; for the LOCAL operators a site sign,
//                                                       generic_ks/spin_taste_ops.c:172-263)
//   meson(x)  = <antiquark(x) | src2(x)>               (su3_dot, :349-363)
//   meson_q[t][p] = sum_{x in slice t} meson(x) ftfact_p(x)      (:364-383)
// with ftfact_p(x) = prod_{d = x,y,z} f(2 pi (x_d - r0_d) p_d / n_d), f = cos, i sin or exp(i .) by the reflection
// parity of the component (ff(), :137-157; table built at :262-283).  On the host that is one pass over two
// propagators plus no_q_momenta complex multiply-adds per site and a no_q_momenta x volume table of phases.
//
// Here: one CTA per chunk of kMesonSites sites of ONE time slice and parity (a time slice is a contiguous range of
// the checkerboard index), a thread keeps the colour dot products of its sites in registers and walks the momenta:
// the phase of a site is the product of three entries of per-direction tables (nmom x (nx + ny + nz) numbers, built
// on the host with the reference's own expression; L1-resident), the CTA's sum for a momentum goes through the warp
// shuffle tree and four shared-memory slots, in a fixed order.  meson_finish_kernel adds the chunks of a slice, also
// in a fixed order: the result is deterministic.  Algorithmic bytes: 96 per site (two colour vectors read once) for
// ANY number of momenta -- HBM-bound up to a few dozen momenta, FP64-issue-bound beyond.
//
// The per-site arithmetic is __host__ __device__ so that host loops can run it against the oracle.
#pragma once
#include "common.cuh"

namespace b200ks {

constexpr int kMesonPerThread = 4;
constexpr int kMesonSites = kBlock * kMesonPerThread;   // sites per CTA
constexpr int kMesonMaxMom = 128;                        // (the reference's MAXQ is 100)

// (-)^[spin . (x - r0)] eps(x - r0)^spin, times the antiquark's (-)^(x+y+z+t - r0): local(), spin_taste_ops.c:245-263.
// h = coordinates relative to r0.  spin < 0: no sign (the caller has applied its operator already).
__host__ __device__ inline double meson_local_sign(int spin, int hx, int hy, int hz, int ht) {
  if (spin < 0) return 1.0;
  const int hp = (hx + hy + hz + ht) & 1;
  int flips = hp;   // antiquark_sign_flip
  if ((spin & 1) && (hp ^ (hx & 1))) flips++;
  if ((spin & 2) && (hp ^ (hy & 1))) flips++;
  if ((spin & 4) && (hp ^ (hz & 1))) flips++;
  if ((spin & 8) && (hp ^ (ht & 1))) flips++;
  return (flips & 1) ? -1.0 : 1.0;
}

// <a|b> at site f
__host__ __device__ inline double2 meson_dot_site(const double2 *a, const double2 *b, size_t stride, int f) {
  double re = 0.0, im = 0.0;
  for (int c = 0; c < 3; c++) {
    const double2 aa = a[(size_t)c * stride + f], bb = b[(size_t)c * stride + f];
    re += aa.x * bb.x + aa.y * bb.y;
    im += aa.x * bb.y - aa.y * bb.x;
  }
  return make_double2(re, im);
}

__host__ __device__ inline double2 meson_cmul(const double2 a, const double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

#ifdef __CUDACC__
struct MesonArg {
  const double2 *anti[2];   // per parity half
  const double2 *quark[2];
  const double2 *tab;       // [nmom][gx + gy + gz] per-direction phase factors, indexed by GLOBAL coordinate
  double2 *partial;         // [L[3]][2][nchunk][nmom]
  int nmom, nchunk, slice_h;   // slice_h = sites per parity in one time slice
  int spin;
  int r0[4];
  Geom g;
};

__global__ void __launch_bounds__(kBlock) meson_mom_kernel(const MesonArg a) {
  __shared__ double2 sm[kBlock / 32][kMesonMaxMom];
  const int t = blockIdx.y, par = blockIdx.z;
  const Geom &g = a.g;
  const int gsum = g.G[0] + g.G[1] + g.G[2];
  double2 z[kMesonPerThread];
  int ox[kMesonPerThread], oy[kMesonPerThread], oz[kMesonPerThread];
#pragma unroll
  for (int k = 0; k < kMesonPerThread; k++) {
    const int in_slice = blockIdx.x * kMesonSites + k * kBlock + threadIdx.x;
    z[k] = make_double2(0.0, 0.0);
    ox[k] = oy[k] = oz[k] = 0;
    if (in_slice < a.slice_h) {
      const int f = t * a.slice_h + in_slice;
      const Coord c = site_coord(g, f, par);
      const int X = c.x + g.origin[0], Y = c.y + g.origin[1], Z = c.z + g.origin[2], T = c.t + g.origin[3];
      const double s = meson_local_sign(a.spin, X - a.r0[0], Y - a.r0[1], Z - a.r0[2], T - a.r0[3]);
      const double2 d = meson_dot_site(a.anti[par], a.quark[par], (size_t)g.stride, f);
      z[k] = make_double2(s * d.x, s * d.y);
      ox[k] = X;
      oy[k] = g.G[0] + Y;
      oz[k] = g.G[0] + g.G[1] + Z;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = 0; p < a.nmom; p++) {
    const double2 *tp = a.tab + (size_t)p * gsum;
    double re = 0.0, im = 0.0;
#pragma unroll
    for (int k = 0; k < kMesonPerThread; k++) {
      // ((1 . f_x) . f_y) . f_z, the reference's order (ks_meson_mom.c:276-280)
      const double2 ph = meson_cmul(meson_cmul(__ldg(tp + ox[k]), __ldg(tp + oy[k])), __ldg(tp + oz[k]));
      re += z[k].x * ph.x - z[k].y * ph.y;
      im += z[k].x * ph.y + z[k].y * ph.x;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      re += __shfl_xor_sync(0xffffffffu, re, o);
      im += __shfl_xor_sync(0xffffffffu, im, o);
    }
    if (lane == 0) sm[warp][p] = make_double2(re, im);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < a.nmom; p += kBlock) {
    double2 s = sm[0][p];
#pragma unroll
    for (int w = 1; w < kBlock / 32; w++) { s.x += sm[w][p].x; s.y += sm[w][p].y; }
    a.partial[(((size_t)t * 2 + par) * a.nchunk + blockIdx.x) * a.nmom + p] = s;
  }
}

// out[t][p] = sum over parity and chunk, fixed order; one thread per (t, p)
__global__ void __launch_bounds__(128) meson_finish_kernel(const double2 *partial, int nt, int nchunk, int nmom, double2 *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nt * nmom) return;
  const int t = i / nmom, p = i - t * nmom;
  double2 s = make_double2(0.0, 0.0);
  for (int k = 0; k < 2 * nchunk; k++) {
    const double2 v = partial[((size_t)t * 2 * nchunk + k) * nmom + p];
    s.x += v.x;
    s.y += v.y;
  }
  out[i] = s;
}
#endif

}  // namespace b200ks
