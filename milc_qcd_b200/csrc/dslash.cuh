// dslash.cuh -- the improved-staggered (fat + Naik) stencil kernel.
//
//   out(x) = sum_mu [ F_mu(x) in(x+mu) + L_mu(x) in(x+3mu)
//                   - F_mu(x-mu)^+ in(x-mu) - L_mu(x-3mu)^+ in(x-3mu) ]   (+ s*w(x))
//
// for all sites x of one parity; `in` lives on the opposite parity.  This is what the
// reference computes in dslash_fn_field_special (generic_ks/dslash_fn_dblstore.c:311-562,
// generic_ks/dslash_fn.c:365-576): forward hops multiply by the link stored at x, backward
// hops by the ADJOINT of the link stored at the neighbour (generic_ks/fn_links_milc.c:132-146,
// 180-194).  KS phases and boundary signs are already inside the links.
//
// One thread per output site.  Per site the kernel streams 16 links (8 of this parity,
// 8 of the other) and gathers 16 neighbour colour vectors, 1146 flop; it is HBM-bound
// (0.48 flop/B in double), so the work is all in the memory system: fully coalesced SoA
// loads, evict-first on the links, default caching on the 16x-reused colour vectors.
//
// Fused epilogues (CG BLAS-1 folded into the stencil, d_congrad5_fn_milc.c:286-308,
// ks_multicg_offset.c:311-320):
//   kXpay : out = D in + s*w            (w on the output parity; s = -4m^2 or shift0)
//   kDot  : additionally red[0] = sum Re<w|out>, red[1] = sum Re<out|r>, red[2] = |out|^2
#pragma once
#include "common.cuh"

namespace b200ks {

template <typename T>
struct DslashArg {
  using T2 = typename Vec2<T>::type;
  Geom g;
  int par;              // parity bit of the OUTPUT sites (0 even, 1 odd)
  const T2 *fat_this;   // links of the output parity  (forward hops)
  const T2 *lng_this;
  const T2 *fat_other;  // links of the input parity   (backward hops, adjoint)
  const T2 *lng_other;
  const T2 *in;         // input colour vector, opposite parity
  T2 *out;              // output colour vector, this parity
  const T2 *w;          // xpay operand (this parity), may alias nothing
  const T2 *r;          // second dot operand (this parity) or nullptr
  T s;                  // xpay coefficient
  ReduceWs ws;
  double *red;          // device result slots for the fused reductions
  const int *stop;      // device flag: nonzero => solver already converged, do nothing
  int site_begin, site_end;  // sub-range of cb sites handled by this launch
  int ghost_mode;       // 0: all hops (single GPU); 1: interior hops only; 2: ghost hops only, accumulate
};

template <typename T, typename T2>
__device__ __forceinline__ void load_vec(const T2 *v, int stride, int i, T2 (&o)[3]) {
#pragma unroll
  for (int c = 0; c < 3; c++) o[c] = ld_keep(v + (size_t)c * stride + i);
}

template <typename T, typename T2>
__device__ __forceinline__ void load_link(const T2 *U, int lstride, int mu, int i, T2 (&o)[9]) {
  const T2 *p = U + (size_t)mu * 9 * lstride + i;
#pragma unroll
  for (int e = 0; e < 9; e++) o[e] = ld_stream(p + (size_t)e * lstride);
}

// acc += U v
template <typename T, typename T2>
__device__ __forceinline__ void mat_vec_add(const T2 (&U)[9], const T2 (&v)[3], T (&acc)[6]) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) cmad<T2, T>(acc[2 * r], acc[2 * r + 1], U[3 * r + k], v[k]);
}
// acc -= U^dagger v     (c_r = sum_k conj(U[k][r]) v_k, libraries/m_amv_4vec.c:48-115)
template <typename T, typename T2>
__device__ __forceinline__ void adj_mat_vec_sub(const T2 (&U)[9], const T2 (&v)[3], T (&acc)[6]) {
  T t[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) cmad_conj<T2, T>(t[2 * r], t[2 * r + 1], U[3 * k + r], v[k]);
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] -= t[k];
}

template <typename T, int D, int kMode>
__device__ __forceinline__ void hop_dir(const DslashArg<T> &a, int idx, const Coord &c, T (&acc)[6]) {
  using T2 = typename Vec2<T>::type;
  const Geom &g = a.g;
  T2 U[9], v[3];
  const int coord = (D == 0) ? c.x : (D == 1) ? c.y : (D == 2) ? c.z : c.t;
  const bool part = (D >= 2) && g.part[D];
#pragma unroll
  for (int hop = 0; hop < 4; hop++) {
    const int h = (hop == 0) ? 1 : (hop == 1) ? 3 : (hop == 2) ? -1 : -3;
    if (kMode != 0) {
      const bool is_ghost = part && (coord + h < 0 || coord + h >= g.L[D]);
      if ((kMode == 1) == is_ghost) continue;
    }
    const bool fwd = hop < 2;
    const bool lng = (hop & 1);
    if (fwd) {
      const int n = neighbor<D, false>(g, idx, c, h);
      load_link<T, T2>(lng ? a.lng_this : a.fat_this, g.lstride, D, idx, U);
      load_vec<T, T2>(a.in, g.stride, n, v);
      mat_vec_add<T, T2>(U, v, acc);
    } else {
      const int n = neighbor<D, false>(g, idx, c, h);
      const int nl = part ? neighbor<D, true>(g, idx, c, h) : n;
      load_link<T, T2>(lng ? a.lng_other : a.fat_other, g.lstride, D, nl, U);
      load_vec<T, T2>(a.in, g.stride, n, v);
      adj_mat_vec_sub<T, T2>(U, v, acc);
    }
  }
}

// kEpi: 0 plain store, 1 xpay, 2 xpay + 3 fused dots.  kMode: ghost_mode (see DslashArg).
template <typename T, int kEpi, int kMode>
__global__ void __launch_bounds__(kBlock) dslash_kernel(const DslashArg<T> a) {
  using T2 = typename Vec2<T>::type;
  if (a.stop != nullptr && *a.stop) return;
  const int idx = a.site_begin + blockIdx.x * kBlock + threadIdx.x;
  const bool active = idx < a.site_end;
  double red[3] = {0, 0, 0};
  if (active) {
    const Coord c = site_coord(a.g, idx, a.par);
    T acc[6] = {0, 0, 0, 0, 0, 0};
    if (kMode == 2) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const T2 o = a.out[(size_t)k * a.g.stride + idx];
        acc[2 * k] = o.x;
        acc[2 * k + 1] = o.y;
      }
    }
    hop_dir<T, 0, kMode>(a, idx, c, acc);
    hop_dir<T, 1, kMode>(a, idx, c, acc);
    hop_dir<T, 2, kMode>(a, idx, c, acc);
    hop_dir<T, 3, kMode>(a, idx, c, acc);
    if (kEpi >= 1) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const T2 wv = a.w[(size_t)k * a.g.stride + idx];
        acc[2 * k] = fma(a.s, wv.x, acc[2 * k]);
        acc[2 * k + 1] = fma(a.s, wv.y, acc[2 * k + 1]);
        if (kEpi == 2) {
          red[0] += (double)wv.x * (double)acc[2 * k] + (double)wv.y * (double)acc[2 * k + 1];
          red[2] += (double)acc[2 * k] * (double)acc[2 * k] + (double)acc[2 * k + 1] * (double)acc[2 * k + 1];
          if (a.r != nullptr) {
            const T2 rv = a.r[(size_t)k * a.g.stride + idx];
            red[1] += (double)rv.x * (double)acc[2 * k] + (double)rv.y * (double)acc[2 * k + 1];
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      T2 o;
      o.x = acc[2 * k];
      o.y = acc[2 * k + 1];
      a.out[(size_t)k * a.g.stride + idx] = o;
    }
  }
  if (kEpi == 2) grid_reduce<3>(red, a.ws, a.red);
}

}  // namespace b200ks
