// mrhs.cuh -- the fat + Naik stencil applied to K right-hand sides in one pass.
//
// The reference's block solver (ks_congrad_block_parity, generic_ks/d_congrad5_fn_milc.c:409-417)
// is a loop over sources, and so is every multi-source caller above it (the stochastic current
// estimators, generic_ks/f_meas_current.c:339,417-419,515-519; the three colours of a point
// source in ks_spectrum).  Each solve streams the same 16 links per site again.  The stencil is
// HBM-bound and 95 % of its bytes are links (2048 of 2144 per site in double), so applying it to
// K colour vectors at once divides the link traffic per right-hand side by K:
//
//     bytes per site per right-hand side = (8*R_fat + 8*R_long) * w / K + 12 w
//     double 18/14: 2144 (K=1) -> 1120 (K=2) -> 779 (K=3) -> 608 (K=4)
//
// One thread per output site as in dslash.cuh; every link is loaded once into registers and
// multiplied into the K neighbour vectors.  The arithmetic per right-hand side is the same
// sequence of fused multiply-adds as dslash_kernel's, and the fused reductions (two-stage:
// block_partials + reduce_finish_kernel) use the same summation tree, so a block solve reproduces
// K single solves bit for bit.
//
// kMode 1 (round 2): partitioned contexts.  One exchange carries the halos of all K inputs (K push kernels into K
// sub-buffers of the exchange's ghost buffer, the last one raises the arrival flags); the launch is ordered
// interior-first like dslash_kernel's, boundary CTAs acquire the flags and read ghost vectors / ghost links where a
// hop crosses a face.
#pragma once
#include "blas.cuh"
#include "dslash.cuh"

namespace b200ks {

// kMaxRhs = 4 right-hand sides per pass (common.cuh): 12 accumulator registers each in double

template <typename T, int K>
struct DslashMArg {
  using T2 = typename Vec2<T>::type;
  Geom g;
  int par;              // parity bit of the OUTPUT sites
  const T2 *fat_this, *lng_this, *fat_other, *lng_other;
  const T2 *in[K];      // input colour vectors (opposite parity)
  T2 *out[K];           // outputs (this parity)
  const T2 *w[K];       // kEpi 2: xpay operands (this parity)
  const T2 *r[K];       // kEpi 2: second dot operands
  const int *stop[K];   // per-right-hand-side stop flag (nullptr: always live)
  T s;
  ReduceWs ws;          // kEpi 2: per-CTA partial sums [CTA][3*K] (slot k: values 3k..3k+2)
  int nsites;
  // kMode 1 (partitioned lattice): ghost zone of every input, site lists, arrival flags as in DslashArg
  const T2 *gin[K];
  const int *sites;
  int n_int, n_ext, nb_int;
  const unsigned long long *halo_flags;
  unsigned long long halo_seq;
  int halo_mask;
  int *halo_err;
  long long halo_timeout;
};

// the four hops of direction D for all K right-hand sides: each link is loaded once
// kPart: the site may have neighbours in the ghost zones (a boundary site of a partitioned lattice)
template <typename T, int D, int K, int kNc, bool kPart>
__device__ __forceinline__ void hop_dir_m(const DslashMArg<T, K> &a, int idx, const Coord &c, T (&acc)[K][6]) {
  using T2 = typename Vec2<T>::type;
  const Geom &g = a.g;
  T2 U[9], v[3];
  const bool part = kPart && (D >= 2) && g.part[D];
#pragma unroll
  for (int hop = 0; hop < 4; hop++) {
    const int h = (hop == 0) ? 1 : (hop == 1) ? 3 : (hop == 2) ? -1 : -3;
    const bool lng = (hop & 1);
    const int n = neighbor<D, false, kPart>(g, idx, c, h);
    if (hop < 2) {
      if (lng) load_long<T, T2, kNc>(a.lng_this, g.lstride, D, idx, U);
      else load_link<T, T2>(a.fat_this, g.lstride, D, idx, U);
    } else {
      const int nl = part ? neighbor<D, true>(g, idx, c, h) : n;   // backward links of ghost sites: tail of the link field
      if (lng) load_long<T, T2, kNc>(a.lng_other, g.lstride, D, nl, U);
      else load_link<T, T2>(a.fat_other, g.lstride, D, nl, U);
    }
    const bool ghost = part && n >= g.Vh;
    const int nv = ghost ? n - g.Vh : n;
    const int vst = ghost ? g.gstride : g.stride;
#pragma unroll
    for (int k = 0; k < K; k++) {
      const T2 *base = ghost ? a.gin[k] : a.in[k];
#pragma unroll
      for (int q = 0; q < 3; q++) v[q] = ld_keep(base + (size_t)q * vst + nv);
      if (hop < 2) mat_vec_add<T, T2>(U, v, acc[k]);
      else adj_mat_vec_sub<T, T2>(U, v, acc[k]);
    }
  }
}

// kEpi 0: out_k = D in_k.   kEpi 2: out_k = D in_k + s*w_k and red_k = {<w|out>, <out|r>, |out|^2}.
// A right-hand side whose stop flag is set is computed (no divergence in the hop loop) but
// neither stored nor reduced; when every flag is set the launch is a no-op.
// (register caps: double 3 CTAs per SM (<= 168), float 5 (<= 102) or 4 at K = 4 (<= 128), so that
// enough link loads are in flight per SM)
#ifndef B200KS_MRHS_MINB_D
#define B200KS_MRHS_MINB_D 3
#endif
#ifndef B200KS_MRHS_MINB_F4
#define B200KS_MRHS_MINB_F4 4
#endif
#ifndef B200KS_MRHS_MINB_F
#define B200KS_MRHS_MINB_F 5
#endif
// the 16 hops of one output site for all K right-hand sides
template <typename T, int K, int kNc, bool kPart>
__device__ __forceinline__ void mrhs_site(const DslashMArg<T, K> &a, int idx, const Coord &c, T (&acc)[K][6]) {
  hop_dir_m<T, 0, K, kNc, kPart>(a, idx, c, acc);
  hop_dir_m<T, 1, K, kNc, kPart>(a, idx, c, acc);
  hop_dir_m<T, 2, K, kNc, kPart>(a, idx, c, acc);
  hop_dir_m<T, 3, K, kNc, kPart>(a, idx, c, acc);
}

template <typename T, int kEpi, int K, int kNc, int kMode>
__global__ void __launch_bounds__(kBlock, sizeof(T) == 8 ? B200KS_MRHS_MINB_D : (K == 4 ? B200KS_MRHS_MINB_F4 : B200KS_MRHS_MINB_F))
dslash_mrhs_kernel(const DslashMArg<T, K> a) {
  using T2 = typename Vec2<T>::type;
  bool live[K];
  bool any = false;
#pragma unroll
  for (int k = 0; k < K; k++) {
    live[k] = (a.stop[k] == nullptr) || (*a.stop[k] == 0);
    any = any || live[k];
  }
  if (!any) return;
  int idx = blockIdx.x * kBlock + threadIdx.x;
  bool active = idx < a.nsites;
  bool bnd = false;
  if (kMode == 1) {   // interior CTAs first, then the boundary-site list behind the arrival flags (dslash_kernel)
    const int b = blockIdx.x;
    bnd = b >= a.nb_int;
    if (bnd) {
      const int k = (b - a.nb_int) * kBlock + threadIdx.x;
      active = k < a.n_ext;
      idx = active ? __ldg(a.sites + k) : 0;
      if (a.halo_flags != nullptr) acquire_halo_cta(a.halo_flags, a.halo_seq, a.halo_mask, a.halo_err, a.halo_timeout);
    } else {
      const int k = b * kBlock + threadIdx.x;
      active = k < a.n_int;
      idx = active ? interior_site(a.g, k) : 0;
    }
  }
  double red[3 * K];
#pragma unroll
  for (int j = 0; j < 3 * K; j++) red[j] = 0;
  if (active) {
    const Coord c = site_coord(a.g, idx, a.par);
    T acc[K][6];
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
      for (int j = 0; j < 6; j++) acc[k][j] = 0;
    if (kMode == 1 && bnd) mrhs_site<T, K, kNc, true>(a, idx, c, acc);
    else mrhs_site<T, K, kNc, false>(a, idx, c, acc);
#pragma unroll
    for (int k = 0; k < K; k++) {
      if (!live[k]) continue;
      if (kEpi == 2) {
        // same per-site arithmetic as dslash_kernel's epilogue (working precision per site,
        // double over sites; d_congrad5_fn_milc.c:210,293)
        T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const T2 wv = a.w[k][(size_t)q * a.g.stride + idx];
          acc[k][2 * q] = fma(a.s, wv.x, acc[k][2 * q]);
          acc[k][2 * q + 1] = fma(a.s, wv.y, acc[k][2 * q + 1]);
          s0 = fma(wv.x, acc[k][2 * q], fma(wv.y, acc[k][2 * q + 1], s0));
          s2 = fma(acc[k][2 * q], acc[k][2 * q], fma(acc[k][2 * q + 1], acc[k][2 * q + 1], s2));
          const T2 rv = a.r[k][(size_t)q * a.g.stride + idx];
          s1 = fma(rv.x, acc[k][2 * q], fma(rv.y, acc[k][2 * q + 1], s1));
        }
        red[3 * k] = s0;
        red[3 * k + 1] = s1;
        red[3 * k + 2] = s2;
      }
#pragma unroll
      for (int q = 0; q < 3; q++) {
        T2 o;
        o.x = acc[k][2 * q];
        o.y = acc[k][2 * q + 1];
        a.out[k][(size_t)q * a.g.stride + idx] = o;
      }
    }
  }
  // two-stage reduction: per-CTA partial sums [3*K] here, reduce_finish_kernel (one CTA per slot,
  // skipping stopped ones) adds them up in grid_reduce's order
  if (kEpi == 2) block_partials<3 * K>(red, a.ws.partials);
}

}  // namespace b200ks
