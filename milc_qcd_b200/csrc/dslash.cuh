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
//   kEpi 1 : out = D in + s*w            (w on the output parity; s = -4m^2 or shift0)
//   kEpi 2 : additionally red[0] = sum Re<w|out>, red[1] = sum Re<out|r>, red[2] = |out|^2
//
// Multi-GPU (lattice split in t, then z; depth-3 ghost zones, the reference's
// D_FN_GATHER13 observation that the 1-hop halo is a subset of the 3-hop halo,
// dslash_fn_dblstore.c:344-416):
//   kMode 0 : single GPU, every hop
//   kMode 1 : interior pass over all sites -- hops that stay on this GPU; the epilogue is
//             applied only to sites with no ghost hop, boundary sites store raw partial sums
//   kMode 2 : exterior pass over the boundary-site list -- adds the hops that read the ghost
//             buffer (filled by the halo exchange that overlapped kMode 1), then the epilogue
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
  const T2 *gin;        // ghost buffer of `in` (multi-GPU), index = neighbour index - Vh
  T2 *out;              // output colour vector, this parity
  const T2 *w;          // xpay operand (this parity)
  const T2 *r;          // second dot operand (this parity) or nullptr
  T s;                  // xpay coefficient
  ReduceWs ws;
  double *red;          // device result slots for the fused reductions
  const int *stop;      // device flag: nonzero => solver already converged, do nothing
  const int *sites;     // kMode 2: list of boundary sites
  int nsites;           // number of threads' worth of work (Vh, or length of `sites`)
};

template <typename T, typename T2>
__device__ __forceinline__ void load_vec(const DslashArg<T> &a, int n, bool part, T2 (&o)[3]) {
  const T2 *base = a.in;
  int st = a.g.stride;
  if (part && n >= a.g.Vh) {
    base = a.gin;
    st = a.g.gstride;
    n -= a.g.Vh;
  }
#pragma unroll
  for (int c = 0; c < 3; c++) o[c] = ld_keep(base + (size_t)c * st + n);
}

template <typename T, typename T2>
__device__ __forceinline__ void load_link(const T2 *U, int lstride, int mu, int i, T2 (&o)[9]) {
  const T2 *p = U + (size_t)mu * 9 * lstride + i;
#pragma unroll
  for (int e = 0; e < 9; e++) o[e] = ld_stream(p + (size_t)e * lstride);
}

// Long (Naik) links are a real scalar times a U(3) matrix (c_naik * W W W with KS/boundary
// signs folded in, generic_ks/fermion_links_hisq_load_milc.c; c3 * U U U for asqtad), so the
// third row is redundant: row3 = f * conj(row1 x row2) with one complex number f = det(V)/s
// per link.  kNc == 7 reads the compressed form {row1, row2, f} = 14 reals instead of 18 and
// rebuilds row 3 in registers (9 complex multiplies against 32 bytes of HBM traffic).
template <typename T, typename T2, int kNc>
__device__ __forceinline__ void load_long(const T2 *U, int lstride, int mu, int i, T2 (&o)[9]) {
  if (kNc == 9) {
    load_link<T, T2>(U, lstride, mu, i, o);
    return;
  }
  const T2 *p = U + (size_t)mu * 7 * lstride + i;
#pragma unroll
  for (int e = 0; e < 6; e++) o[e] = ld_stream(p + (size_t)e * lstride);
  const T2 f = ld_stream(p + (size_t)6 * lstride);
  reconstruct_row3<T, T2>(o, f);
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

template <typename T, int D, int kMode, int kNc>
__device__ __forceinline__ void hop_dir(const DslashArg<T> &a, int idx, const Coord &c, T (&acc)[6]) {
  using T2 = typename Vec2<T>::type;
  const Geom &g = a.g;
  T2 U[9], v[3];
  const int coord = (D == 0) ? c.x : (D == 1) ? c.y : (D == 2) ? c.z : c.t;
  const bool part = (kMode != 0) && (D >= 2) && g.part[D];
  if (kMode == 2 && !part) return;
#pragma unroll
  for (int hop = 0; hop < 4; hop++) {
    const int h = (hop == 0) ? 1 : (hop == 1) ? 3 : (hop == 2) ? -1 : -3;
    if (kMode != 0) {
      const bool is_ghost = part && (coord + h < 0 || coord + h >= g.L[D]);
      if ((kMode == 1) == is_ghost) continue;
    }
    const bool lng = (hop & 1);
    const int n = neighbor<D, false>(g, idx, c, h);
    if (hop < 2) {
      if (lng) load_long<T, T2, kNc>(a.lng_this, g.lstride, D, idx, U);
      else load_link<T, T2>(a.fat_this, g.lstride, D, idx, U);
      load_vec<T, T2>(a, n, part, v);
      mat_vec_add<T, T2>(U, v, acc);
    } else {
      const int nl = part ? neighbor<D, true>(g, idx, c, h) : n;
      if (lng) load_long<T, T2, kNc>(a.lng_other, g.lstride, D, nl, U);
      else load_link<T, T2>(a.fat_other, g.lstride, D, nl, U);
      load_vec<T, T2>(a, n, part, v);
      adj_mat_vec_sub<T, T2>(U, v, acc);
    }
  }
}

__device__ __forceinline__ bool is_boundary(const Geom &g, const Coord &c) {
  bool b = false;
  if (g.part[2]) b = b || (c.z < 3) || (c.z >= g.L[2] - 3);
  if (g.part[3]) b = b || (c.t < 3) || (c.t >= g.L[3] - 3);
  return b;
}

// kEpi: 0 plain store, 1 xpay, 2 xpay + 3 fused dots.  kMode: see the header comment.
// kNc: complex numbers stored per long link (9 = full matrix, 7 = two rows + U(3) factor).
template <typename T, int kEpi, int kMode, int kNc>
__global__ void __launch_bounds__(kBlock) dslash_kernel(const DslashArg<T> a) {
  using T2 = typename Vec2<T>::type;
  if (a.stop != nullptr && *a.stop) return;
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const bool active = k < a.nsites;
  double red[3] = {0, 0, 0};
  if (active) {
    const int idx = (kMode == 2) ? a.sites[k] : k;
    const Coord c = site_coord(a.g, idx, a.par);
    T acc[6] = {0, 0, 0, 0, 0, 0};
    if (kMode == 2) {
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const T2 o = a.out[(size_t)q * a.g.stride + idx];
        acc[2 * q] = o.x;
        acc[2 * q + 1] = o.y;
      }
    }
    hop_dir<T, 0, kMode, kNc>(a, idx, c, acc);
    hop_dir<T, 1, kMode, kNc>(a, idx, c, acc);
    hop_dir<T, 2, kMode, kNc>(a, idx, c, acc);
    hop_dir<T, 3, kMode, kNc>(a, idx, c, acc);
    const bool do_epi = (kMode != 1) || !is_boundary(a.g, c);
    if (kEpi >= 1 && do_epi) {
#pragma unroll
      for (int q = 0; q < 3; q++) {
        const T2 wv = a.w[(size_t)q * a.g.stride + idx];
        acc[2 * q] = fma(a.s, wv.x, acc[2 * q]);
        acc[2 * q + 1] = fma(a.s, wv.y, acc[2 * q + 1]);
        if (kEpi == 2) {
          red[0] += (double)wv.x * (double)acc[2 * q] + (double)wv.y * (double)acc[2 * q + 1];
          red[2] += (double)acc[2 * q] * (double)acc[2 * q] + (double)acc[2 * q + 1] * (double)acc[2 * q + 1];
          if (a.r != nullptr) {
            const T2 rv = a.r[(size_t)q * a.g.stride + idx];
            red[1] += (double)rv.x * (double)acc[2 * q] + (double)rv.y * (double)acc[2 * q + 1];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 3; q++) {
      T2 o;
      o.x = acc[2 * q];
      o.y = acc[2 * q + 1];
      a.out[(size_t)q * a.g.stride + idx] = o;
    }
  }
  if (kEpi == 2) grid_reduce<3>(red, a.ws, a.red);
}

// z faces are strided in memory (3 z-slices for every t): gather them into a contiguous
// send buffer [side][colour][slice][t][y][xh]; t faces are already contiguous per colour
// and are sent straight from the field.
template <typename T>
__global__ void __launch_bounds__(kBlock)
pack_zface_kernel(typename Vec2<T>::type *buf, const typename Vec2<T>::type *v, Geom g) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  const int face3 = 3 * g.faceh[2];
  if (k >= 2 * face3) return;
  const int side = k / face3;           // 0: low slices z=0..2 (go to the backward neighbour's "ahead" ghost)
  const int r = k - side * face3;       // 1: high slices z=L-3..L-1
  const int slice = r / g.faceh[2];
  const int within = r - slice * g.faceh[2];   // t*S2 + (y*Lxh + xh)
  const int S2 = g.Lxh * g.L[1];
  const int t = within / S2, r2 = within - t * S2;
  const int z = side ? g.L[2] - 3 + slice : slice;
  const int idx = (t * g.L[2] + z) * S2 + r2;
#pragma unroll
  for (int c = 0; c < 3; c++) buf[(size_t)(side * 3 + c) * face3 + r] = v[(size_t)c * g.stride + idx];
}

}  // namespace b200ks
