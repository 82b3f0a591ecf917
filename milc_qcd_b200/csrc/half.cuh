// half.cuh -- 16-bit fixed-point storage for the inner Krylov iteration of the mixed-precision CG.
//
// What MILC's MAX_MIXED build asks of the GPU seam (inv_args.mixed_precision = 2,
// generic_ks/d_congrad5_fn_gpu.c:104-111): the outer solution and every true residual stay
// double, the inner iteration streams 16-bit links and a 16-bit search direction.  All
// arithmetic is fp32; only the STORAGE of the two stencil operands is 16 bit:
//
//   colour vector : 4 planes of 32-bit words per parity half, plane c < 3 = colour c as two
//                   offset-binary u16 (re | im << 16), plane 3 = the site's scale (float bits):
//                   value = (q - 32768) * scale / 32767.                       16 B/site
//   fat link      : 9 words per direction, one scale for the whole field        36 B/link
//   long link     : compressed form (rows 1,2 + U(3) factor f) = 7 words with one scale for
//                   the rows and one for f, or 9 words when the links are not compressible
//                                                                               28 B/link
//   => 8*36 + 8*28 + 16 + 16 = 544 B per output site per stencil (float: 1072, double: 2144).
//
// At 544 B/site the kernel is no longer purely HBM-bound: instruction issue matters (first
// version: 2687 instructions per site, 64 % issue utilisation at 52 % of DRAM peak, ncu).  Three
// things keep the count down:
//   * u16 -> float costs one PRMT and one packed FADD per PAIR of numbers (no conversion-pipe
//     instruction): the 16 bits are dropped into the mantissa of 2^23 and the offset is
//     subtracted exactly, giving integer-valued floats; the scales (one per site for vectors,
//     one per field for links) are applied once per hop to the 6 numbers of the product;
//   * sm_100 packed fp32 (FFMA2/FADD2/FMUL2): hops are processed two at a time -- the same hop
//     type (fat/long, forward/backward) in two directions, one per lane of a float2 -- so the
//     whole matrix-vector product, the row rebuild and the conversions issue half as often;
//   * tiled layout: words are stored [site/32][plane][site%32], so the planes of one site are
//     compile-time offsets (plane*128 B) from one address instead of one 64-bit address
//     computation per plane, and a warp still reads whole 128-byte lines.
#pragma once
#include "dslash.cuh"

namespace b200ks {

constexpr float kHalfBias = 8388608.0f + 32768.0f;   // 2^23 + offset-binary zero

// word index of plane 0 of site i in a tiled array with npl planes; plane p is at +32*p
__device__ __forceinline__ unsigned tile_base(int i, int npl) { return ((unsigned)i >> 5) * (unsigned)(npl * 32) + ((unsigned)i & 31u); }
inline size_t tile_words(size_t nsites, int npl) { return (nsites + 31) / 32 * 32 * (size_t)npl; }

__device__ __forceinline__ float magic_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); }
__device__ __forceinline__ float magic_hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }
__device__ __forceinline__ float2 unpack_raw(uint32_t w) {   // (q - 32768) as exact floats
  return make_float2(magic_lo(w) - kHalfBias, magic_hi(w) - kHalfBias);
}
__device__ __forceinline__ float2 unpack_h(uint32_t w, float k) {   // k = scale/32767
  const float2 r = unpack_raw(w);
  return make_float2(r.x * k, r.y * k);
}
__device__ __forceinline__ uint32_t pack_h(float re, float im, float inv) {   // inv = 32767/scale
  int a = __float2int_rn(re * inv), b = __float2int_rn(im * inv);
  a = max(-32767, min(32767, a)) + 32768;
  b = max(-32767, min(32767, b)) + 32768;
  return (uint32_t)a | ((uint32_t)b << 16);
}

// quantise one site's colour vector (6 floats) into its 4 planes (tiled layout)
__device__ __forceinline__ void store_vec_h(uint32_t *v, int i, const float (&x)[6]) {
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < 6; k++) m = fmaxf(m, fabsf(x[k]));
  const float inv = m > 0.f ? 32767.0f / m : 0.f;
  uint32_t *p = v + tile_base(i, 4);
#pragma unroll
  for (int c = 0; c < 3; c++) p[32 * c] = pack_h(x[2 * c], x[2 * c + 1], inv);
  p[96] = __float_as_uint(m);
}
__device__ __forceinline__ void load_vec_h(const uint32_t *v, int i, float2 (&o)[3]) {
  const uint32_t *p = v + tile_base(i, 4);
  const float k = __uint_as_float(__ldg(p + 96)) * (1.0f / 32767.0f);
#pragma unroll
  for (int c = 0; c < 3; c++) o[c] = unpack_h(__ldg(p + 32 * c), k);
}

struct HalfLinks {
  const uint32_t *fat[2];   // [parity] 36 planes, tiled
  const uint32_t *lng[2];   // [parity] 4*nc planes, tiled
  float fat_k, lng_k, f_k;  // scale/32767 of fat components, long rows, long factor
};

struct DslashHArg {
  Geom g;
  int par;
  HalfLinks L;
  const uint32_t *in;    // half colour vector, opposite parity
  const uint32_t *gin;   // ghost buffer (4 planes, tiled)
  uint32_t *out_h;       // kEpi 0: half output
  float2 *out_f;         // kEpi 2: float output (A p, consumed by the float update kernel)
  const uint32_t *w_h;   // kEpi 2: xpay operand (the half search direction, output parity)
  const float2 *r;       // kEpi 2: residual (float)
  float s;
  ReduceWs ws;
  double *red;
  const int *stop;
  const int *sites;
  int nsites, n_int, n_ext, nb_int, blk0;
  const unsigned long long *halo_flags;
  unsigned long long halo_seq;
  int halo_mask;
  int *halo_err;
  long long halo_timeout;
};

// ---- packed fp32 helpers (sm_100 FFMA2 / FADD2 / FMUL2) -------------------------------------------
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }

// Two hops at once: the same hop type in directions DA (lane x) and DB (lane y).
//   kBack false: acc += U(x) v(x + h)          kBack true: acc -= U(x - h)^dagger v(x - h)
// acc[j] holds real number j (re0, im0, re1, ...) of the two lanes' partial sums.
template <int DA, int DB, bool kBack, bool kLong, int kMode, int kNc>
__device__ __forceinline__ void hop_pair_h(const DslashHArg &a, int idx, const Coord &c, bool bnd, float2 (&acc)[6]) {
  const Geom &g = a.g;
  constexpr int nc = kLong ? kNc : 9;
  const int h = (kLong ? 3 : 1) * (kBack ? -1 : 1);
  const bool partA = (kMode == 1) && (DA >= 2) && bnd && g.part[DA];
  const bool partB = (kMode == 1) && (DB >= 2) && bnd && g.part[DB];
  const int nA = neighbor<DA, false>(g, idx, c, h), nB = neighbor<DB, false>(g, idx, c, h);
  const uint32_t *vA = (partA && nA >= g.Vh) ? a.gin + tile_base(nA - g.Vh, 4) : a.in + tile_base(nA, 4);
  const uint32_t *vB = (partB && nB >= g.Vh) ? a.gin + tile_base(nB - g.Vh, 4) : a.in + tile_base(nB, 4);
  const int lA = !kBack ? idx : partA ? neighbor<DA, true>(g, idx, c, h) : nA;
  const int lB = !kBack ? idx : partB ? neighbor<DB, true>(g, idx, c, h) : nB;
  const uint32_t *links = kLong ? a.L.lng[kBack ? a.par ^ 1 : a.par] : a.L.fat[kBack ? a.par ^ 1 : a.par];
#ifdef B200KS_PROBE_NOLINKLOAD   // diagnostic build: every site reads the links of tile 0 (L1-resident) => compute time only
  const uint32_t *uA = links + (threadIdx.x & 31) + DA * nc * 32;
  const uint32_t *uB = links + (threadIdx.x & 31) + DB * nc * 32;
#else
  const uint32_t *uA = links + tile_base(lA, 4 * nc) + DA * nc * 32;
  const uint32_t *uB = links + tile_base(lB, 4 * nc) + DB * nc * 32;
#endif

  const float2 mB = make_float2(-kHalfBias, -kHalfBias), pB = make_float2(kHalfBias, kHalfBias);
  const float2 neg1 = make_float2(-1.f, -1.f);
  // neighbour vectors; the negated copy supplies the minus sign of the complex product
  float2 vre[3], vim[3], vneg[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint32_t wa = __ldg(vA + 32 * k), wb = __ldg(vB + 32 * k);
    const float2 lo = make_float2(magic_lo(wa), magic_lo(wb)), hi = make_float2(magic_hi(wa), magic_hi(wb));
    vre[k] = padd(lo, mB);
    vim[k] = padd(hi, mB);
    vneg[k] = kBack ? pfma(lo, neg1, pB) : pfma(hi, neg1, pB);   // backward: -re, forward: -im
  }
  const float kvA = __uint_as_float(__ldg(vA + 96)), kvB = __uint_as_float(__ldg(vB + 96));
  // links as integer-valued floats
  float2 ure[9], uim[9];
  constexpr int nload = (nc == 7) ? 6 : 9;
#pragma unroll
  for (int e = 0; e < nload; e++) {
    const uint32_t wa = __ldcs(uA + 32 * e), wb = __ldcs(uB + 32 * e);
    ure[e] = padd(make_float2(magic_lo(wa), magic_lo(wb)), mB);
    uim[e] = padd(make_float2(magic_hi(wa), magic_hi(wb)), mB);
  }
#ifdef B200KS_PROBE_NOMATH       // diagnostic build: loads only (same addresses, same order) => memory time only
  {
    float2 sum = make_float2(0.f, 0.f);
#pragma unroll
    for (int e = 0; e < nload; e++) sum = padd(sum, padd(ure[e], uim[e]));
#pragma unroll
    for (int k = 0; k < 3; k++) sum = padd(sum, padd(vre[k], vim[k]));
    if (nc == 7) {
      const uint32_t wa = __ldcs(uA + 32 * 6), wb = __ldcs(uB + 32 * 6);
      sum = padd(sum, make_float2(magic_lo(wa), magic_lo(wb)));
    }
    acc[0] = pfma(make_float2(kvA, kvB), sum, acc[0]);
    return;
  }
#endif
  if (nc == 7) {  // row3 = f * conj(row1 x row2), f brought to the rows' integer grid
    const uint32_t wa = __ldcs(uA + 32 * 6), wb = __ldcs(uB + 32 * 6);
    const float fk = a.L.f_k * a.L.lng_k;
    const float2 fk2 = make_float2(fk, fk);
    const float2 fre = pmul(padd(make_float2(magic_lo(wa), magic_lo(wb)), mB), fk2);
    const float2 fim = pmul(padd(make_float2(magic_hi(wa), magic_hi(wb)), mB), fk2);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      // d = U[k1]*U[3+k2] - U[k2]*U[3+k1] ;  c = conj(d)
      float2 dre = pmul(ure[k1], ure[3 + k2]);
      dre = pfma(pmul(uim[k1], neg1), uim[3 + k2], dre);
      dre = pfma(pmul(ure[k2], neg1), ure[3 + k1], dre);
      dre = pfma(uim[k2], uim[3 + k1], dre);
      float2 dim = pmul(ure[k1], uim[3 + k2]);
      dim = pfma(uim[k1], ure[3 + k2], dim);
      dim = pfma(pmul(ure[k2], neg1), uim[3 + k1], dim);
      dim = pfma(pmul(uim[k2], neg1), ure[3 + k1], dim);
      // f * (dre - i dim) = (fre*dre + fim*dim) + i (fim*dre - fre*dim)
      ure[6 + k] = pfma(fre, dre, pmul(fim, dim));
      uim[6 + k] = pfma(fim, dre, pmul(pmul(fre, neg1), dim));
    }
  }
  // t = U v (forward) or U^dagger v (backward)
  const float2 zero = make_float2(0.f, 0.f);
  float2 t[6] = {zero, zero, zero, zero, zero, zero};
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int e = kBack ? 3 * k + r : 3 * r + k;
      if (!kBack) {   // (ure + i uim)(vre + i vim)
        t[2 * r] = pfma(ure[e], vre[k], t[2 * r]);
        t[2 * r] = pfma(uim[e], vneg[k], t[2 * r]);
        t[2 * r + 1] = pfma(ure[e], vim[k], t[2 * r + 1]);
        t[2 * r + 1] = pfma(uim[e], vre[k], t[2 * r + 1]);
      } else {        // (ure - i uim)(vre + i vim)
        t[2 * r] = pfma(ure[e], vre[k], t[2 * r]);
        t[2 * r] = pfma(uim[e], vim[k], t[2 * r]);
        t[2 * r + 1] = pfma(ure[e], vim[k], t[2 * r + 1]);
        t[2 * r + 1] = pfma(uim[e], vneg[k], t[2 * r + 1]);
      }
    }
  const float ku = (kLong ? a.L.lng_k : a.L.fat_k) * (kBack ? -1.0f / 32767.0f : 1.0f / 32767.0f);
  const float2 sc = make_float2(ku * kvA, ku * kvB);
#pragma unroll
  for (int j = 0; j < 6; j++) acc[j] = pfma(sc, t[j], acc[j]);
}

// kEpi 0: out_h = D in.   kEpi 2: out_f = D in + s*w_h, red = {<w|out>, <out|r>, |out|^2}.
#ifndef B200KS_HALF_MINBLOCKS
#define B200KS_HALF_MINBLOCKS 6   // CTAs per SM the register allocation is held to (80 registers)
#endif
template <int kEpi, int kMode, int kNc>
__global__ void __launch_bounds__(kBlock, B200KS_HALF_MINBLOCKS) dslash_half_kernel(const DslashHArg a) {
  if (a.stop != nullptr && *a.stop) return;
  int k = blockIdx.x * kBlock + threadIdx.x;
  bool active = k < a.nsites;
  bool bnd = false;
  if (kMode == 1) {
    const int b = blockIdx.x + a.blk0;
    bnd = b >= a.nb_int;
    if (bnd) {
      k = (b - a.nb_int) * kBlock + threadIdx.x;
      active = k < a.n_ext;
      if (a.halo_flags != nullptr) {
        if (threadIdx.x == 0) acquire_halo(a.halo_flags, a.halo_seq, a.halo_mask, a.halo_err, a.halo_timeout);
        __syncthreads();
      }
    } else {
      k = b * kBlock + threadIdx.x;
      active = k < a.n_int;
    }
  }
  double red[3] = {0, 0, 0};
  if (active) {
    const int idx = (kMode == 0) ? k : bnd ? a.sites[k] : interior_site(a.g, k);
    const Coord c = site_coord(a.g, idx, a.par);
    const float2 zero = make_float2(0.f, 0.f);
    float2 acc2[6] = {zero, zero, zero, zero, zero, zero};
    hop_pair_h<0, 1, false, false, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<2, 3, false, false, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<0, 1, false, true, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<2, 3, false, true, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<0, 1, true, false, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<2, 3, true, false, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<0, 1, true, true, kMode, kNc>(a, idx, c, bnd, acc2);
    hop_pair_h<2, 3, true, true, kMode, kNc>(a, idx, c, bnd, acc2);
    float acc[6];
#pragma unroll
    for (int j = 0; j < 6; j++) acc[j] = acc2[j].x + acc2[j].y;
    if (kEpi == 0) {
      store_vec_h(a.out_h, idx, acc);
    } else {
      float2 w[3];
      load_vec_h(a.w_h, idx, w);
      // per-site sums in fp32 (what MILC's single-precision su3_rdot / magsq_su3vec return),
      // accumulated over sites in double (d_congrad5_fn_milc.c:210,293)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < 3; q++) {
        acc[2 * q] = fmaf(a.s, w[q].x, acc[2 * q]);
        acc[2 * q + 1] = fmaf(a.s, w[q].y, acc[2 * q + 1]);
        const float2 rv = a.r[(size_t)q * a.g.stride + idx];
        s0 = fmaf(w[q].x, acc[2 * q], fmaf(w[q].y, acc[2 * q + 1], s0));
        s1 = fmaf(rv.x, acc[2 * q], fmaf(rv.y, acc[2 * q + 1], s1));
        s2 = fmaf(acc[2 * q], acc[2 * q], fmaf(acc[2 * q + 1], acc[2 * q + 1], s2));
        a.out_f[(size_t)q * a.g.stride + idx] = make_float2(acc[2 * q], acc[2 * q + 1]);
      }
      red[0] = s0;
      red[1] = s1;
      red[2] = s2;
    }
  }
  if (kEpi == 2) {   // two-stage (reduce_finish_kernel follows) unless the NCCL-halo path asks for in-kernel sums
    if (kMode == 0 || a.red == nullptr) block_partials<3>(red, a.ws.partials);
    else grid_reduce<3>(red, a.ws, a.red);
  }
}

// x += a p ; r += a ttt ; p = r + b p (re-quantised) ; sum |r|^2.   x, r, ttt float; p half.
__global__ void __launch_bounds__(kBlock)
cg_update_half_kernel(float2 *x, float2 *r, uint32_t *p_h, const float2 *ttt, int stride, int n, CgState *st, ReduceWs ws,
                      int fuse_scalar) {
  if (st->stop) return;
  const double rsq = st->rsq, oldrsq = st->upd[0];
  const double pkp = st->red[0], c_tr = st->red[1], c_tt = st->red[2];
  const float a = (float)(-rsq / pkp);
  const double rsq_new = oldrsq + 2.0 * (double)a * c_tr + (double)a * (double)a * c_tt;
  const float bb = (float)(rsq_new / oldrsq);
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[2] = {0, 0};
  if (i < n) {
    float2 pv[3];
    load_vec_h(p_h, i, pv);
    float pn[6];
    float rn = 0.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      float2 xv = x[o], rv = r[o];
      const float2 tv = ttt[o];
      xv.x = fmaf(a, pv[c].x, xv.x);
      xv.y = fmaf(a, pv[c].y, xv.y);
      rv.x = fmaf(a, tv.x, rv.x);
      rv.y = fmaf(a, tv.y, rv.y);
      pn[2 * c] = fmaf(bb, pv[c].x, rv.x);
      pn[2 * c + 1] = fmaf(bb, pv[c].y, rv.y);
      x[o] = xv;
      r[o] = rv;
      rn = fmaf(rv.x, rv.x, fmaf(rv.y, rv.y, rn));
    }
    store_vec_h(p_h, i, pn);
    s[0] = rn;
  }
  if (fuse_scalar & 8) {   // two-stage: reduce_finish_kernel sums the partials and advances the recurrence
    block_partials<2>(s, ws.partials);
    return;
  }
  const bool last = grid_reduce<2>(s, ws, st->upd_next);
  if (last && fuse_scalar && threadIdx.x == 0) cg_scalar_step(st, (fuse_scalar >> 1) & 1, (fuse_scalar >> 2) & 1);
}

// Reliable update with a half search direction (see mixed_reliable_kernel in blas.cuh).
__global__ void __launch_bounds__(kBlock)
mixed_reliable_half_kernel(const double2 *b, const double2 *ttt, float2 *r_lo, uint32_t *p_h, int stride, int n, int first,
                           ReduceWs ws, double *out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[2] = {0, 0};
  if (i < n) {
    float2 pv[3];
    if (!first) load_vec_h(p_h, i, pv);
    float pn[6];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      const double2 bv = b[o], tv = ttt[o];
      const double rx = bv.x + tv.x, ry = bv.y + tv.y;
      const float2 ro = r_lo[o];
      const float2 rn = make_float2((float)rx, (float)ry);
      r_lo[o] = rn;
      pn[2 * c] = first ? rn.x : pv[c].x + (rn.x - ro.x);
      pn[2 * c + 1] = first ? rn.y : pv[c].y + (rn.y - ro.y);
      s[0] += rx * rx + ry * ry;
    }
    store_vec_h(p_h, i, pn);
  }
  grid_reduce<2>(s, ws, out);
}

// double <-> 16-bit colour vectors (tests, timing probes)
__global__ void __launch_bounds__(kBlock) vec_d2h_kernel(uint32_t *h, const double2 *d, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  float x[6];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const double2 v = d[(size_t)c * stride + i];
    x[2 * c] = (float)v.x;
    x[2 * c + 1] = (float)v.y;
  }
  store_vec_h(h, i, x);
}
__global__ void __launch_bounds__(kBlock) vec_h2d_kernel(double2 *d, const uint32_t *h, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  float2 v[3];
  load_vec_h(h, i, v);
#pragma unroll
  for (int c = 0; c < 3; c++) d[(size_t)c * stride + i] = make_double2((double)v[c].x, (double)v[c].y);
}

// ---- link quantisation ---------------------------------------------------------------------------
// max |component| of planes [p0, p1) of every link direction (ncomp planes per direction)
template <typename T>
__global__ void __launch_bounds__(kBlock)
link_absmax_kernel(const typename Vec2<T>::type *U, int lstride, int n, int nc, int p0, int p1, unsigned *out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  float m = 0.f;
  if (i < n)
    for (int mu = 0; mu < 4; mu++)
      for (int e = p0; e < p1; e++) {
        const auto v = U[(size_t)(mu * nc + e) * lstride + i];
        m = fmaxf(m, fmaxf(fabsf((float)v.x), fabsf((float)v.y)));
      }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
quantize_link_kernel(uint32_t *dst, const typename Vec2<T>::type *U, int lstride, int n, int nc, float inv_rows, float inv_f) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  for (int mu = 0; mu < 4; mu++)
    for (int e = 0; e < nc; e++) {
      const auto v = U[(size_t)(mu * nc + e) * lstride + i];
      const float inv = (nc == 7 && e == 6) ? inv_f : inv_rows;
      dst[tile_base(i, 4 * nc) + (mu * nc + e) * 32] = pack_h((float)v.x, (float)v.y, inv);
    }
}

}  // namespace b200ks
