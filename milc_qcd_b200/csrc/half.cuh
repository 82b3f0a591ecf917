// half.cuh -- 16-bit fixed-point storage for the inner Krylov iteration of the mixed-precision CG.
//
// What MILC's MAX_MIXED build asks of the GPU seam (inv_args.mixed_precision = 2,
// generic_ks/d_congrad5_fn_gpu.c:104-111): the outer solution and every true residual stay
// double, the inner iteration streams 16-bit links and a 16-bit search direction.  All
// arithmetic is fp32; only the STORAGE of the two stencil operands is 16 bit:
//
//   colour vector : one 16-byte word per site, {c0, c1, c2, scale}: colour c as two offset-binary
//                   u16 (re | im << 16), scale = the site's max |component| (float bits):
//                   value = (q - 32768) * scale / 32767.                         16 B/site
//   fat link      : 9 words, one scale for the whole field                       36 B/link
//   long link     : compressed form (rows 1,2 + U(3) factor f) = 7 words with one scale for
//                   the rows and one for f, or 9 words when the links are not compressible
//                                                                                28 B/link
//   => 8*36 + 8*28 + 16 + 16 = 544 B per output site per stencil (float: 1072, double: 2144).
//
// At 544 B/site the kernel is no longer purely HBM-bound: instruction issue matters (round 1:
// 1714 warp instructions per 32 sites at 72 % of the measured HBM peak, of which 193 were 4-byte
// loads with their own address arithmetic).  What keeps the count down:
//   * u16 -> float costs one PRMT and one packed FADD per PAIR of numbers (no conversion-pipe
//     instruction): the 16 bits are dropped into the mantissa of 2^23 and the offset is
//     subtracted exactly, giving integer-valued floats; the scales (one per site for vectors,
//     one per field for links) are applied once per hop to the 6 numbers of the product;
//   * sm_100 packed fp32 (FFMA2/FADD2/FMUL2): hops are processed two at a time -- the same hop
//     type (fat/long, forward/backward) in two directions, one per lane of a float2 -- so the
//     whole matrix-vector product, the row rebuild and the conversions issue half as often;
//   * 16-BYTE LOADS ONLY (round 2).  The links of one output site are ONE record of 32 (36)
//     16-byte words, stored [site/32][word][site%32] so that a warp reads 512 contiguous bytes per
//     load.  The record holds everything the site multiplies with: its 4 forward fat links, its 4
//     forward long links, and -- stored a second time, already adjointed and negated, like the
//     reference's own fatback/lngback arrays (generic_ks/fn_links_milc.c:114-199) -- the 4 + 4
//     links of its backward neighbours.  Backward hops therefore need no neighbour index for the
//     link, no ghost links and no adjoint code path, and all 16 products are acc += U v.  Inside a
//     hop type the words are ordered [direction pair][element][direction of the pair], which is the
//     order the packed arithmetic consumes them in.  A colour vector is one 16-byte word.
//     Per site: 36 + 16 loads of 16 bytes instead of 193 of 4 bytes; traffic unchanged
//     (the 16-bit link copy doubles to 2 x 0.27 GB at 32^3x64, of 180 GB).
#pragma once
#include "dslash.cuh"

namespace b200ks {

constexpr float kHalfBias = 8388608.0f + 32768.0f;   // 2^23 + offset-binary zero

// 16-byte words per site record: [fwd fat 9][fwd long nq][back fat 9][back long nq], nq = 7 | 9
__host__ __device__ constexpr int half_long_quads(int nc) { return nc == 7 ? 7 : 9; }
__host__ __device__ constexpr int half_record_quads(int nc) { return 2 * (9 + half_long_quads(nc)); }
inline size_t half_link_bytes(size_t nsites, int nc) { return (nsites + 31) / 32 * 32 * (size_t)half_record_quads(nc) * 16; }
// first 16-byte word of site i's record; word q is at + 32*q
__device__ __forceinline__ size_t record_base(int i, int nq) { return (size_t)((unsigned)i >> 5) * (unsigned)(nq * 32) + ((unsigned)i & 31u); }

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

// quantise one site's colour vector (6 floats) into its 16-byte word
__device__ __forceinline__ uint4 store_vec_h(uint4 *v, int i, const float (&x)[6]) {
  float m = 0.f;
#pragma unroll
  for (int k = 0; k < 6; k++) m = fmaxf(m, fabsf(x[k]));
  const float inv = m > 0.f ? 32767.0f / m : 0.f;
  uint4 o;
  o.x = pack_h(x[0], x[1], inv);
  o.y = pack_h(x[2], x[3], inv);
  o.z = pack_h(x[4], x[5], inv);
  o.w = __float_as_uint(m);
  v[i] = o;
  return o;
}
__device__ __forceinline__ void load_vec_h(const uint4 *v, int i, float2 (&o)[3]) {
  const uint4 w = __ldg(v + i);
  const float k = __uint_as_float(w.w) * (1.0f / 32767.0f);
  o[0] = unpack_h(w.x, k);
  o[1] = unpack_h(w.y, k);
  o[2] = unpack_h(w.z, k);
}

struct HalfLinks {
  const uint4 *rec[2];      // [parity] site records (see the header comment)
  float fat_k, lng_k, f_k;  // scale/32767 of fat components, long-link components, long factor
};

struct DslashHArg {
  Geom g;
  int par;
  HalfLinks L;
  const uint4 *in;       // half colour vector, opposite parity
  const uint4 *gin;      // ghost buffer, index = neighbour index - Vh
  uint4 *out_h;          // kEpi 0: half output
  float2 *out_f;         // kEpi 2: float output (A p, consumed by the float update kernel)
  const uint4 *w_h;      // kEpi 2: xpay operand (the half search direction, output parity)
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
  PushArg push;          // kEpi 0, kMode 1, push_on: the boundary sites of the OUTPUT go straight to the neighbours'
  int push_on;           // ghost buffers (fused halo push, dslash.cuh push_site_h), for the stencil that reads it next
  HaloRaise raise;       // arrival flags of the halo the kernel BEFORE this one pushed (its input's, common.cuh)
};

// ---- packed fp32 helpers (sm_100 FFMA2 / FADD2 / FMUL2) -------------------------------------------
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
// negation folds into the operand modifiers of the consuming FFMA2 / FADD2 (no instruction)
__device__ __forceinline__ float2 pneg(float2 a) { return make_float2(-a.x, -a.y); }

__device__ __forceinline__ uint4 ld_rec(const uint4 *p) { return __ldcs(p); }   // links: streamed once per stencil

// Two hops at once: hop type (kLong, kBack) in directions 2P (lane x) and 2P+1 (lane y).
// acc += U v in both lanes; the record already holds -U(x-h)^dagger for the backward types.
// acc[j] holds real number j (re0, im0, re1, ...) of the two lanes' partial sums.
// kPart: the site may have neighbours in the ghost zones (a boundary site of a partitioned lattice);
// interior sites and unpartitioned lattices take the plain periodic index arithmetic.
template <int P, bool kBack, bool kLong, bool kPart, int kNc>
__device__ __forceinline__ void hop_pair_h(const DslashHArg &a, const uint4 *rec, int idx, const Coord &c, float2 (&acc)[6]) {
  const Geom &g = a.g;
  constexpr int DA = 2 * P, DB = 2 * P + 1;
  constexpr int nc = kLong ? kNc : 9;
  constexpr int nql = half_long_quads(kNc);
  constexpr int q0 = (kBack ? 9 + nql : 0) + (kLong ? 9 : 0);   // first 16-byte word of this hop type
  constexpr int w0 = P * 2 * nc;                                // first 4-byte word of this pair inside it
  constexpr int qa = w0 / 4, qb = (w0 + 2 * nc - 1) / 4;        // 16-byte words to load
  constexpr int off = w0 - 4 * qa;
  const int h = (kLong ? 3 : 1) * (kBack ? -1 : 1);
  const bool partA = kPart && (DA >= 2) && g.part[DA];
  const bool partB = kPart && (DB >= 2) && g.part[DB];
  const int nA = neighbor<DA, false, kPart>(g, idx, c, h), nB = neighbor<DB, false, kPart>(g, idx, c, h);
  const uint4 *vA = (partA && nA >= g.Vh) ? a.gin + (nA - g.Vh) : a.in + nA;
  const uint4 *vB = (partB && nB >= g.Vh) ? a.gin + (nB - g.Vh) : a.in + nB;
  const uint4 va = __ldg(vA), vb = __ldg(vB);
  uint32_t w[4 * (qb - qa + 1)];
#pragma unroll
  for (int q = qa; q <= qb; q++) {
#ifdef B200KS_PROBE_NOLINKLOAD   // diagnostic build: every site reads record 0 (L1-resident) => compute time only
    const uint4 x = __ldg(a.L.rec[a.par] + (threadIdx.x & 31) + 32 * (q0 + q));
#else
    const uint4 x = ld_rec(rec + 32 * (q0 + q));
#endif
    w[4 * (q - qa) + 0] = x.x;
    w[4 * (q - qa) + 1] = x.y;
    w[4 * (q - qa) + 2] = x.z;
    w[4 * (q - qa) + 3] = x.w;
  }
  const float2 mB = make_float2(-kHalfBias, -kHalfBias);
  // neighbour vectors
  float2 vre[3], vim[3];
  {
    const uint32_t wa[3] = {va.x, va.y, va.z}, wb[3] = {vb.x, vb.y, vb.z};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float2 lo = make_float2(magic_lo(wa[k]), magic_lo(wb[k])), hi = make_float2(magic_hi(wa[k]), magic_hi(wb[k]));
      vre[k] = padd(lo, mB);
      vim[k] = padd(hi, mB);
    }
  }
  const float kvA = __uint_as_float(va.w), kvB = __uint_as_float(vb.w);
  // links as integer-valued floats
  float2 ure[9], uim[9];
  constexpr int nload = (nc == 7) ? 6 : 9;
#pragma unroll
  for (int e = 0; e < nload; e++) {
    const uint32_t wa = w[off + 2 * e], wb = w[off + 2 * e + 1];
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
    if (nc == 7) sum = padd(sum, make_float2(magic_lo(w[off + 12]), magic_lo(w[off + 13])));
    acc[0] = pfma(make_float2(kvA, kvB), sum, acc[0]);
    return;
  }
#endif
  if (nc == 7) {  // row3 = f * conj(row1 x row2), f brought to the rows' integer grid
    const uint32_t wa = w[off + 12], wb = w[off + 13];
    const float fk = a.L.f_k * a.L.lng_k;
    const float2 fk2 = make_float2(fk, fk);
    const float2 fre = pmul(padd(make_float2(magic_lo(wa), magic_lo(wb)), mB), fk2);
    const float2 fim = pmul(padd(make_float2(magic_hi(wa), magic_hi(wb)), mB), fk2);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      // d = U[k1]*U[3+k2] - U[k2]*U[3+k1] ;  c = conj(d)
      float2 dre = pmul(ure[k1], ure[3 + k2]);
      dre = pfma(pneg(uim[k1]), uim[3 + k2], dre);
      dre = pfma(pneg(ure[k2]), ure[3 + k1], dre);
      dre = pfma(uim[k2], uim[3 + k1], dre);
      float2 dim = pmul(ure[k1], uim[3 + k2]);
      dim = pfma(uim[k1], ure[3 + k2], dim);
      dim = pfma(pneg(ure[k2]), uim[3 + k1], dim);
      dim = pfma(pneg(uim[k2]), ure[3 + k1], dim);
      // f * (dre - i dim) = (fre*dre + fim*dim) + i (fim*dre - fre*dim)
      ure[6 + k] = pfma(fre, dre, pmul(fim, dim));
      uim[6 + k] = pfma(fim, dre, pmul(pneg(fre), dim));
    }
  }
  // t = U v : (ure + i uim)(vre + i vim)
  const float2 zero = make_float2(0.f, 0.f);
  float2 t[6] = {zero, zero, zero, zero, zero, zero};
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int e = 3 * r + k;
      t[2 * r] = pfma(ure[e], vre[k], t[2 * r]);
      t[2 * r] = pfma(pneg(uim[e]), vim[k], t[2 * r]);
      t[2 * r + 1] = pfma(ure[e], vim[k], t[2 * r + 1]);
      t[2 * r + 1] = pfma(uim[e], vre[k], t[2 * r + 1]);
    }
  const float ku = (kLong ? a.L.lng_k : a.L.fat_k) * (1.0f / 32767.0f);
  const float2 sc = make_float2(ku * kvA, ku * kvB);
#pragma unroll
  for (int j = 0; j < 6; j++) acc[j] = pfma(sc, t[j], acc[j]);
}

// one output site: the 16 hops and the epilogue
template <int kEpi, bool kPart, int kNc>
__device__ __forceinline__ void half_site(const DslashHArg &a, int idx, double (&red)[3]) {
  const Coord c = site_coord(a.g, idx, a.par);
  const uint4 *rec = a.L.rec[a.par] + record_base(idx, half_record_quads(kNc));
  const float2 zero = make_float2(0.f, 0.f);
  float2 acc2[6] = {zero, zero, zero, zero, zero, zero};
  hop_pair_h<0, false, false, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<1, false, false, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<0, false, true, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<1, false, true, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<0, true, false, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<1, true, false, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<0, true, true, kPart, kNc>(a, rec, idx, c, acc2);
  hop_pair_h<1, true, true, kPart, kNc>(a, rec, idx, c, acc2);
  float acc[6];
#pragma unroll
  for (int j = 0; j < 6; j++) acc[j] = acc2[j].x + acc2[j].y;
  if (kEpi == 0) {
    const uint4 o = store_vec_h(a.out_h, idx, acc);
    if (kPart && a.push_on) push_site_h(a.push, a.g, idx, c.z, c.t, o);   // (fused halo push of the output)
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

// kEpi 0: out_h = D in.   kEpi 2: out_f = D in + s*w_h, red = {<w|out>, <out|r>, |out|^2}.
#ifndef B200KS_HALF_MINBLOCKS
#define B200KS_HALF_MINBLOCKS 5   // CTAs per SM the register allocation is held to (96 registers: no spills;
                                  // measured 0.0862 ms per launch at 32^3x64 against 0.0925 with 6 and 0.0974 with 7)
#endif
template <int kEpi, int kMode, int kNc>
__global__ void __launch_bounds__(kBlock, B200KS_HALF_MINBLOCKS) dslash_half_kernel(const DslashHArg a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.stop != nullptr && *a.stop) return;
  if (kMode == 1 && blockIdx.x == 0 && threadIdx.x == 0) raise_halo_flags(a.raise);   // (before this CTA waits for anything)
  int k = blockIdx.x * kBlock + threadIdx.x;
  bool active = k < a.nsites;
  bool bnd = false;
  int bsite = 0;
  if (kMode == 1) {
    const int b = blockIdx.x + a.blk0;
    bnd = b >= a.nb_int;
    if (bnd) {
      k = (b - a.nb_int) * kBlock + threadIdx.x;
      active = k < a.n_ext;
      if (active) bsite = __ldg(a.sites + k);   // (in flight while the flags are polled)
      if (a.halo_flags != nullptr) acquire_halo_cta(a.halo_flags, a.halo_seq, a.halo_mask, a.halo_err, a.halo_timeout);
    } else {
      k = b * kBlock + threadIdx.x;
      active = k < a.n_int;
    }
  }
  double red[3] = {0, 0, 0};
  if (active) {
    const int idx = (kMode == 0) ? k : bnd ? bsite : interior_site(a.g, k);
    // interior sites of a partitioned lattice never leave the local volume: they run the same instruction stream
    // as an unpartitioned lattice (the ghost-index arithmetic of the boundary sites cost every site ~10 %)
    if (kMode == 1 && bnd) half_site<kEpi, true, kNc>(a, idx, red);
    else half_site<kEpi, false, kNc>(a, idx, red);
  }
  // (the boundary CTAs hold exactly the sites the neighbours need: half_site pushed them; the arrival flags are raised
  // by the next kernel on the stream, HaloRaise)
  if (kEpi == 2) {   // two-stage (reduce_finish_kernel follows) unless the NCCL-halo path asks for in-kernel sums
    if (kMode == 0 || a.red == nullptr) block_partials<3>(red, a.ws.partials);
    else grid_reduce<3>(red, a.ws, a.red);
  }
}

// fused halo push of the update kernel's new search direction (on != 0: partitioned context, peer-to-peer halos): the
// next stencil's exchange is under way before this kernel has ended (dslash.cuh push_site_h)
struct HalfPush {
  PushArg a;
  Geom g;
  int par;   // parity bit of the sites this launch updates
  int on;
};

// x += a p ; r += a ttt ; p = r + b p (re-quantised) ; sum |r|^2.   x, r, ttt float; p half.
__global__ void __launch_bounds__(kBlock)
cg_update_half_kernel(float2 *x, float2 *r, uint4 *p_h, const float2 *ttt, int stride, int n, CgState *st, ReduceWs ws,
                      int fuse_scalar, const HalfPush hp) {
  pdl_launch_dependents();
  pdl_wait();
  if (st->stop) return;
  const double rsq = st->rsq, oldrsq = st->upd[0];
  const double pkp = st->red[0], c_tr = st->red[1], c_tt = st->red[2];
  const float a = (float)(-rsq / pkp);
  const double rsq_new = oldrsq + 2.0 * (double)a * c_tr + (double)a * (double)a * c_tt;
  const float bb = (float)(rsq_new / oldrsq);
  const int i = blockIdx.x * kBlock + threadIdx.x;
  const double2 *xrel = (fuse_scalar & 2) ? st->xrel : nullptr;   // Fermilab relative residue wanted (CgState::xrel)
  double s[2] = {0, 0};
  if (i < n) {
    float2 pv[3];
    load_vec_h(p_h, i, pv);
    float pn[6];
    float rn = 0.f;
    double xn2 = 0;
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
      if (xrel != nullptr) {
        const double2 xd = xrel[o];
        const double tx = xd.x + (double)xv.x, ty = xd.y + (double)xv.y;
        xn2 += tx * tx + ty * ty;
      }
    }
    const uint4 o = store_vec_h(p_h, i, pn);
    if (hp.on) {
      const Coord c = site_coord(hp.g, i, hp.par);
      push_site_h(hp.a, hp.g, i, c.z, c.t, o);
    }
    s[0] = rn;
    if (xrel != nullptr) s[1] = (xn2 == 0) ? 1.0 : (double)rn / xn2;
  }
  // (the finish kernel that follows raises the arrival flags of the push above, FinishArg::raise)
  if (fuse_scalar & 8) {   // two-stage: reduce_finish_kernel sums the partials and advances the recurrence
    block_partials<2>(s, ws.partials);
    return;
  }
  const bool last = grid_reduce<2>(s, ws, st->upd_next);
  if (last && fuse_scalar && threadIdx.x == 0) cg_scalar_step(st, (fuse_scalar >> 1) & 1, (fuse_scalar >> 2) & 1);
}

// Reliable update with a half search direction (see mixed_reliable_kernel in blas.cuh).
__global__ void __launch_bounds__(kBlock)
mixed_reliable_half_kernel(const double2 *b, const double2 *ttt, float2 *r_lo, uint4 *p_h, int stride, int n, int first,
                           ReduceWs ws, double *out, const double2 *xrel) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[2] = {0, 0};
  if (i < n) {
    float2 pv[3];
    if (!first) load_vec_h(p_h, i, pv);
    float pn[6];
    double num = 0, den = 0;
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
      if (xrel != nullptr) {
        const double2 xv = xrel[o];
        num += rx * rx + ry * ry;
        den += xv.x * xv.x + xv.y * xv.y;
      }
    }
    store_vec_h(p_h, i, pn);
    if (xrel != nullptr) s[1] = (den == 0) ? 1.0 : num / den;
  }
  grid_reduce<2>(s, ws, out);
}

// double <-> 16-bit colour vectors (tests, timing probes)
__global__ void __launch_bounds__(kBlock) vec_d2h_kernel(uint4 *h, const double2 *d, int stride, int n) {
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
__global__ void __launch_bounds__(kBlock) vec_h2d_kernel(double2 *d, const uint4 *h, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  float2 v[3];
  load_vec_h(h, i, v);
#pragma unroll
  for (int c = 0; c < 3; c++) d[(size_t)c * stride + i] = make_double2((double)v[c].x, (double)v[c].y);
}

// ---- building the site records from the master links (double or single SoA, ghost tails filled) ----
// Loads link mu of `field` at site index i (nc = 7: rows 1, 2 and the factor; row 3 rebuilt).
template <typename T, int kNc>
__device__ __forceinline__ void load_master(const typename Vec2<T>::type *field, int lstride, int mu, int i,
                                            typename Vec2<T>::type (&U)[9]) {
  using T2 = typename Vec2<T>::type;
  const T2 *p = field + (size_t)mu * kNc * lstride + i;
  if (kNc == 9) {
#pragma unroll
    for (int e = 0; e < 9; e++) U[e] = p[(size_t)e * lstride];
  } else {
#pragma unroll
    for (int e = 0; e < 6; e++) U[e] = p[(size_t)e * lstride];
    reconstruct_row3<T, T2>(U, p[(size_t)6 * lstride]);
  }
}
// B = -U^dagger
template <typename T2>
__device__ __forceinline__ void neg_adjoint(const T2 (&U)[9], T2 (&B)[9]) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int k = 0; k < 3; k++) {
      B[3 * r + k].x = -U[3 * k + r].x;
      B[3 * r + k].y = U[3 * k + r].y;
    }
}

// max |component| of the fat links (out[0]), of the FULL long links incl. the rebuilt third row
// (out[1]; the backward copies store columns as rows) and of the U(3) factors (out[2]; a link's
// factor and its adjoint's have the same modulus), over n sites (ghost tails included)
template <typename T, int kNc>
__global__ void __launch_bounds__(kBlock)
half_absmax_kernel(const typename Vec2<T>::type *fat, const typename Vec2<T>::type *lng, int lstride, int n, unsigned *out) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  float m[3] = {0.f, 0.f, 0.f};
  if (i < n)
    for (int mu = 0; mu < 4; mu++) {
      T2 U[9];
      load_master<T, 9>(fat, lstride, mu, i, U);
#pragma unroll
      for (int e = 0; e < 9; e++) m[0] = fmaxf(m[0], fmaxf(fabsf((float)U[e].x), fabsf((float)U[e].y)));
      load_master<T, kNc>(lng, lstride, mu, i, U);
#pragma unroll
      for (int e = 0; e < 9; e++) m[1] = fmaxf(m[1], fmaxf(fabsf((float)U[e].x), fabsf((float)U[e].y)));
      if (kNc == 7) {
        const T2 f = lng[(size_t)(mu * 7 + 6) * lstride + i];
        m[2] = fmaxf(m[2], fmaxf(fabsf((float)f.x), fabsf((float)f.y)));
      }
    }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[k] = fmaxf(m[k], __shfl_xor_sync(0xffffffffu, m[k], o));
    if ((threadIdx.x & 31) == 0 && m[k] > 0.f) atomicMax(out + k, __float_as_uint(m[k]));
  }
}

// 4-byte word `wi` of a hop type that starts at 16-byte word q0 of the record at `out`
__device__ __forceinline__ void put_word(uint32_t *out, int q0, int wi, uint32_t word) {
  out[(size_t)128 * (q0 + (wi >> 2)) + (wi & 3)] = word;
}
// one link of direction mu into the hop type at q0 (kLong: a long link, stored with kNc words)
template <typename T, int kNc, bool kLong>
__device__ __forceinline__ void put_link(uint32_t *out, int q0, int mu, const typename Vec2<T>::type (&U)[9], float inv, float inv_f) {
  constexpr int nc = kLong ? kNc : 9;
  const int base = (mu >> 1) * 2 * nc + (mu & 1);   // word of element e: base + 2*e
  if (!kLong || kNc == 9) {
#pragma unroll
    for (int e = 0; e < 9; e++) put_word(out, q0, base + 2 * e, pack_h((float)U[e].x, (float)U[e].y, inv));
  } else {
    double fx, fy, dev;
    long_factor<T>(U, fx, fy, dev);
#pragma unroll
    for (int e = 0; e < 6; e++) put_word(out, q0, base + 2 * e, pack_h((float)U[e].x, (float)U[e].y, inv));
    put_word(out, q0, base + 12, pack_h((float)fx, (float)fy, inv_f));
  }
}

// One thread per output site of parity `par`: writes the site's record.  this_* = master links of the
// output parity (forward hops), other_* = of the opposite parity (backward hops: the link stored at
// the backward neighbour, or in the backward-ghost tail of the field on a partitioned lattice).
template <typename T, int kNc, int MU>
__device__ __forceinline__ void record_dir(uint32_t *out, const typename Vec2<T>::type *fat_this, const typename Vec2<T>::type *lng_this,
                                           const typename Vec2<T>::type *fat_other, const typename Vec2<T>::type *lng_other,
                                           const Geom &g, int idx, const Coord &c, float inv_fat, float inv_lng, float inv_f) {
  using T2 = typename Vec2<T>::type;
  constexpr int nql = half_long_quads(kNc);
  T2 U[9], B[9];
  load_master<T, 9>(fat_this, g.lstride, MU, idx, U);
  put_link<T, kNc, false>(out, 0, MU, U, inv_fat, inv_f);
  load_master<T, kNc>(lng_this, g.lstride, MU, idx, U);
  put_link<T, kNc, true>(out, 9, MU, U, inv_lng, inv_f);
  load_master<T, 9>(fat_other, g.lstride, MU, neighbor<MU, true>(g, idx, c, -1), U);
  neg_adjoint<T2>(U, B);
  put_link<T, kNc, false>(out, 9 + nql, MU, B, inv_fat, inv_f);
  load_master<T, kNc>(lng_other, g.lstride, MU, neighbor<MU, true>(g, idx, c, -3), U);
  neg_adjoint<T2>(U, B);
  put_link<T, kNc, true>(out, 18 + nql, MU, B, inv_lng, inv_f);
}
template <typename T, int kNc>
__global__ void __launch_bounds__(kBlock)
half_records_kernel(uint4 *rec, const typename Vec2<T>::type *fat_this, const typename Vec2<T>::type *lng_this,
                    const typename Vec2<T>::type *fat_other, const typename Vec2<T>::type *lng_other, const Geom g, int par,
                    float inv_fat, float inv_lng, float inv_f) {
  const int idx = blockIdx.x * kBlock + threadIdx.x;
  if (idx >= g.Vh) return;
  const Coord c = site_coord(g, idx, par);
  uint32_t *out = (uint32_t *)(rec + record_base(idx, half_record_quads(kNc)));
  record_dir<T, kNc, 0>(out, fat_this, lng_this, fat_other, lng_other, g, idx, c, inv_fat, inv_lng, inv_f);
  record_dir<T, kNc, 1>(out, fat_this, lng_this, fat_other, lng_other, g, idx, c, inv_fat, inv_lng, inv_f);
  record_dir<T, kNc, 2>(out, fat_this, lng_this, fat_other, lng_other, g, idx, c, inv_fat, inv_lng, inv_f);
  record_dir<T, kNc, 3>(out, fat_this, lng_this, fat_other, lng_other, g, idx, c, inv_fat, inv_lng, inv_f);
}

}  // namespace b200ks
