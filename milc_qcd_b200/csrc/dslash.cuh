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
// dslash_fn_dblstore.c:344-416).  Every site is computed exactly once, with all 16 hops and
// its epilogue:
//   kMode 0 : single GPU, all sites
//   kMode 1 : partitioned lattice.  The grid is ordered: the first nb_int CTAs take the
//             interior sites (at least 3 slices away from every partitioned face; no hop
//             leaves the GPU), the remaining CTAs take the boundary sites (site list), whose
//             hops across a partitioned face read the ghost buffer and the backward-ghost tail
//             of the link fields.  With peer-to-peer halos it is ONE launch: CTAs are issued in
//             index order, so the interior CTAs stream while the neighbours' push kernels fill
//             the ghost buffer, and each boundary CTA first acquires the arrival flags
//             (halo_flags/halo_seq).  With NCCL halos the two ranges are two launches (blk0
//             selects the range) separated by a stream event.
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
  const int *sites;     // kMode 1: list of boundary sites
  int nsites;           // kMode 0: Vh
  int n_int, n_ext;     // kMode 1: interior / boundary site counts
  int nb_int;           // kMode 1: CTAs covering the interior sites
  int blk0;             // kMode 1: logical index of this launch's first CTA
  const unsigned long long *halo_flags;  // kMode 1, peer-to-peer: arrival flags to acquire (nullptr: none)
  unsigned long long halo_seq;
  int halo_mask;
  int *halo_err;
  long long halo_timeout;
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

// kPart: the site may have neighbours in the ghost zones (boundary site of a partitioned lattice)
template <typename T, int D, bool kPart, int kNc>
__device__ __forceinline__ void hop_dir(const DslashArg<T> &a, int idx, const Coord &c, T (&acc)[6]) {
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

// k-th interior site (kMode 1): z and t run over [3, L-3) in partitioned directions
__device__ __forceinline__ int interior_site(const Geom &g, int k) {
  const int q = fast_div(k, g.dS2), r = k - q * g.S2;
  const int tq = fast_div(q, g.dZi);
  const int z = q - tq * g.zi + (g.part[2] ? 3 : 0);
  const int t = tq + (g.part[3] ? 3 : 0);
  return (t * g.L[2] + z) * g.S2 + r;
}

// kEpi: 0 plain store, 1 xpay, 2 xpay + 3 fused dots.  kMode: see the header comment.
// kNc: complex numbers stored per long link (9 = full matrix, 7 = two rows + U(3) factor).
// Spin (bounded) until every expected face of exchange `seq` has arrived; flags live in this
// GPU's memory and are raised by the neighbours' push kernels (push_halo_kernel).
__device__ __forceinline__ void acquire_halo(const unsigned long long *flags, unsigned long long seq, int mask, int *err,
                                             long long max_cycles) {
  const long long t0 = clock64();
  for (int f = 0; f < 4; f++) {
    if (!((mask >> f) & 1)) continue;
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + f) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > max_cycles) {
        *err = 1 + f;
        break;
      }
    }
  }
}

// The same for a whole CTA (every thread calls it; ends with a CTA barrier).  B200KS_ACQ selects how the flags are
// polled (A/B builds, profiles/variants):
//   0  thread 0 polls the faces one after the other with ld.acquire.sys (round 1)
//   1  lanes 0..3 poll one face each, ld.acquire.sys
//   2  lanes 0..3 poll one face each with ld.relaxed.sys; no fence: the ghost words are read after the barrier, through
//      L1 lines that cannot be stale (nothing reads a ghost buffer before its flag has been seen, and the L1 starts
//      every kernel empty), and the data reached this GPU's L2 before the flag did (the writer's release)
//   3  as 2 + fence.acq_rel.sys in the polling lanes (the formal acquire pattern)
// Measured with the GPU as its own neighbour at the 8-GPU local volume (profiles/run_r02k.sh): 16-bit stencil 0.1675 /
// 0.1644 / 0.1495 / 0.1595 ms for 0 / 1 / 2 / 3 (unpartitioned: 0.131-0.137).  ld.acquire.sys and the system fence
// compile to CCTL.IVALL -- every boundary CTA emptied its SM's L1, which the resident CTAs' neighbour-vector reuse lives in.
#ifndef B200KS_ACQ
#define B200KS_ACQ 2
#endif
__device__ __forceinline__ void acquire_halo_cta(const unsigned long long *flags, unsigned long long seq, int mask, int *err,
                                                 long long max_cycles) {
  if (B200KS_ACQ == 0) {
    if (threadIdx.x == 0) acquire_halo(flags, seq, mask, err, max_cycles);
  } else if (threadIdx.x < 4 && ((mask >> threadIdx.x) & 1)) {
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      if (B200KS_ACQ == 1) asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
      else asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + threadIdx.x) : "memory");
      if (v >= seq) break;
      if (clock64() - t0 > max_cycles) {
        *err = 1 + threadIdx.x;
        break;
      }
    }
    if (B200KS_ACQ == 3) asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncthreads();
}

// one output site: the 16 hops and the epilogue
template <typename T, int kEpi, bool kPart, int kNc>
__device__ __forceinline__ void dslash_site(const DslashArg<T> &a, int idx, double (&red)[3]) {
  using T2 = typename Vec2<T>::type;
  const Coord c = site_coord(a.g, idx, a.par);
  T acc[6] = {0, 0, 0, 0, 0, 0};
  hop_dir<T, 0, kPart, kNc>(a, idx, c, acc);
  hop_dir<T, 1, kPart, kNc>(a, idx, c, acc);
  hop_dir<T, 2, kPart, kNc>(a, idx, c, acc);
  hop_dir<T, 3, kPart, kNc>(a, idx, c, acc);
  if (kEpi >= 1) {
    // per-site sums in the working precision (MILC's su3_rdot / magsq_su3vec return Real),
    // accumulated over sites in double (d_congrad5_fn_milc.c:210,293)
    T s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
    for (int q = 0; q < 3; q++) {
      const T2 wv = a.w[(size_t)q * a.g.stride + idx];
      acc[2 * q] = fma(a.s, wv.x, acc[2 * q]);
      acc[2 * q + 1] = fma(a.s, wv.y, acc[2 * q + 1]);
      if (kEpi == 2) {
        s0 = fma(wv.x, acc[2 * q], fma(wv.y, acc[2 * q + 1], s0));
        s2 = fma(acc[2 * q], acc[2 * q], fma(acc[2 * q + 1], acc[2 * q + 1], s2));
        if (a.r != nullptr) {
          const T2 rv = a.r[(size_t)q * a.g.stride + idx];
          s1 = fma(rv.x, acc[2 * q], fma(rv.y, acc[2 * q + 1], s1));
        }
      }
    }
    red[0] = s0;
    red[1] = s1;
    red[2] = s2;
  }
#pragma unroll
  for (int q = 0; q < 3; q++) {
    T2 o;
    o.x = acc[2 * q];
    o.y = acc[2 * q + 1];
    a.out[(size_t)q * a.g.stride + idx] = o;
  }
}

// (register caps keep 5 CTAs per SM in double and 8 in float whatever epilogue is compiled in)
template <typename T, int kEpi, int kMode, int kNc>
__global__ void __launch_bounds__(kBlock, sizeof(T) == 8 ? 5 : 8) dslash_kernel(const DslashArg<T> a) {
  using T2 = typename Vec2<T>::type;
  pdl_launch_dependents();
  pdl_wait();
  if (a.stop != nullptr && *a.stop) return;
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
    // interior sites of a partitioned lattice run the unpartitioned instruction stream
    if (kMode == 1 && bnd) dslash_site<T, kEpi, true, kNc>(a, idx, red);
    else dslash_site<T, kEpi, false, kNc>(a, idx, red);
  }
  if (kEpi == 2) {   // two-stage (reduce_finish_kernel follows) unless the NCCL-halo path asks for in-kernel sums
    if (kMode == 0 || a.red == nullptr) block_partials<3>(red, a.ws.partials);
    else grid_reduce<3>(red, a.ws, a.red);
  }
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

// ---- peer-to-peer halo push (comm.cuh P2P) ------------------------------------------------------
// Stores this rank's 3 low and 3 high slices of every partitioned direction into the
// neighbours' ghost buffers (mapped peer memory, NVLink stores), then the last CTA raises the
// arrival flags.  Low slices land in the backward neighbour's "ahead" zone, high slices in the
// forward neighbour's "behind" zone; all ranks share one local geometry, so offsets are ours.
struct PushArg {
  void *dst[2][2];        // [d-2][0: our low slices | 1: our high slices] destination ghost buffer
  unsigned long long *flag[2][2];   // arrival flag to raise at that destination
  unsigned long long seq;
  unsigned *ticket;
  const int *stop;        // solver stop flag: a stopped solver's stencil launches are no-ops, so is this
};

// ---- fused halo push (16-bit colour vectors) -------------------------------------------------------
// The kernel that PRODUCES a vector stores the words of its boundary sites straight into the neighbours' ghost
// buffers (NVLink stores from the registers that hold the result) instead of leaving them to push_halo_kernel, which
// would have to wait for the producer to end, be launched behind an event on another stream and read the words back:
// the transfer overlaps the producer's own work tile by tile, and the kernel launched next raises the arrival flags
// (common.cuh HaloRaise: no fence and no ticket in the producer).
// push_site_h: site idx (local coordinates z, t) of the produced parity.
// Offsets as in push_halo_kernel: low slices land in the backward neighbour's "ahead" zone, high slices in the
// forward neighbour's "behind" zone; with extents below 6 a site can be in both bands.
__device__ __forceinline__ void push_site_h(const PushArg &a, const Geom &g, int idx, int z, int t, const uint4 w) {
  if (g.part[3]) {
    const int within = idx - t * g.faceh[3];
    if (t < 3) {
      ((uint4 *)a.dst[1][0])[(g.ghost[3][1] - g.Vh) + t * g.faceh[3] + within] = w;
    }
    if (t >= g.L[3] - 3) {
      ((uint4 *)a.dst[1][1])[(g.ghost[3][0] - g.Vh) + (t - (g.L[3] - 3)) * g.faceh[3] + within] = w;
    }
  }
  if (g.part[2]) {
    const int within = idx - (t * g.L[2] + z) * g.S2 + t * g.S2;   // t*S2 + (y*Lxh + xh)
    if (z < 3) {
      ((uint4 *)a.dst[0][0])[(g.ghost[2][1] - g.Vh) + z * g.faceh[2] + within] = w;
    }
    if (z >= g.L[2] - 3) {
      ((uint4 *)a.dst[0][1])[(g.ghost[2][0] - g.Vh) + (z - (g.L[2] - 3)) * g.faceh[2] + within] = w;
    }
  }
}
constexpr int kPushBlock = 256;

// A few long-lived CTAs (one per SM at most) with a grid-stride loop: remote stores are
// fire-and-forget, so a CTA keeps issuing until its share is done and pays the NVLink round
// trip once, at the fence.  (Many short CTAs would each sit in an SM slot for a round trip,
// starving the concurrent interior stencil CTAs of slots.)
// E = element type of one plane (double2 / float2: 3 colour planes, plane stride = field
// stride; uint32_t with kTile: the 4 planes of a 16-bit colour vector in the tiled layout
// [site/32][plane][site%32] of half.cuh), NP = planes.
template <typename E, int NP, bool kTile>
__global__ void __launch_bounds__(kPushBlock)
push_halo_kernel(const PushArg a, const E *v, const Geom g) {
  if (a.stop != nullptr && *a.stop) return;
  const int nz = g.part[2] ? 6 * g.faceh[2] : 0;
  const int nt = g.part[3] ? 6 * g.faceh[3] : 0;
  const int S2 = g.Lxh * g.L[1];
  for (int k = blockIdx.x * kPushBlock + threadIdx.x; k < nz + nt; k += gridDim.x * kPushBlock) {
    const int d = (k < nz) ? 2 : 3;
    const int r = (k < nz) ? k : k - nz;
    const int face3 = 3 * g.faceh[d];
    const int side = r / face3;
    const int r2 = r - side * face3;
    const int slice = r2 / g.faceh[d];
    const int within = r2 - slice * g.faceh[d];
    int idx;
    if (d == 3) {
      idx = (side ? g.L[3] - 3 + slice : slice) * g.faceh[3] + within;
    } else {
      const int t = within / S2, rr = within - t * S2;
      idx = (t * g.L[2] + (side ? g.L[2] - 3 + slice : slice)) * S2 + rr;
    }
    const int off = (g.ghost[d][side ? 0 : 1] - g.Vh) + slice * g.faceh[d] + within;
    E *dst = (E *)a.dst[d - 2][side];
    E x[NP];
    if (kTile) {
      const E *sp = v + (((unsigned)idx >> 5) * (unsigned)(NP * 32) + ((unsigned)idx & 31u));
      E *dp = dst + (((unsigned)off >> 5) * (unsigned)(NP * 32) + ((unsigned)off & 31u));
#pragma unroll
      for (int c = 0; c < NP; c++) x[c] = sp[32 * c];
#pragma unroll
      for (int c = 0; c < NP; c++) dp[32 * c] = x[c];
    } else {
#pragma unroll
      for (int c = 0; c < NP; c++) x[c] = v[(size_t)c * g.stride + idx];
#pragma unroll
      for (int c = 0; c < NP; c++) dst[(size_t)c * g.gstride + off] = x[c];
    }
  }
  // release: the CTA barrier orders every thread's stores before thread 0's system fence
  // (fence cumulativity), which orders them before the ticket and, in the last CTA, the flags
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    __threadfence_system();
    last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  if (threadIdx.x < 4) {
    const int d2 = threadIdx.x >> 1, side = threadIdx.x & 1;
    __threadfence_system();
    // atomicMax, not a store: flags only ever move forward, whatever order two exchanges'
    // updates reach the neighbour in
    if (g.part[d2 + 2] && a.flag[d2][side] != nullptr) atomicMax_system(a.flag[d2][side], a.seq);
  }
  if (threadIdx.x == 0) *a.ticket = 0;
}

}  // namespace b200ks
