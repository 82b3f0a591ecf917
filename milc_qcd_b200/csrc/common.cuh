// common.cuh -- device-side building blocks shared by all kernels of libb200ks.
//
// Device data layout:
//   * checkerboarded: every field is split into an EVEN and an ODD half; inside a half a
//     site is addressed by cb = lex >> 1 with lex = x + Lx*(y + Ly*(z + Lz*t)), which is
//     exactly MILC's index inside a parity block (generic/layout_hyper_prime.c:509-520),
//     so host<->device re-layout is a pure AoS<->SoA transpose, no permutation.
//   * structure of arrays of complex pairs: colour vector element c of site cb lives at
//     v[c*stride + cb] (T2 = double2/float2), link element e = 3*row+col of direction mu
//     at U[(mu*9 + e)*stride + cb].  Consecutive threads = consecutive cb = consecutive
//     16-byte (double2) words: every warp load is one fully coalesced 512-byte request.
//   * ghost zones (multi-GPU) are appended after the Vh interior sites of each half.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace b200ks {

constexpr int kBlock = 128;           // threads per CTA for site kernels
constexpr int kMaxShifts = 32;
constexpr int kMaxRhs = 4;            // right-hand sides per pass of the block solver (mrhs.cuh)

// Launch order of FULL-lattice kernels (both parities, site f = parity*Vh + cb) whose sites gather from
// neighbours: the link construction's staple passes and the fermion force.  In the plain order (thread i = site i)
// the grid walks all even sites, then all odd ones.  Every hop changes parity, so the even half of the grid streams
// the odd halves of its input fields from HBM (as neighbours) and the even halves (as the sites themselves and their
// two-hop neighbours), and the odd half of the grid streams all of it again, a gigabyte later, long after the L2 has
// lost it: an input field that is read at both parities costs 2 x 144 B per site instead of 144.  Interleaved, a CTA
// takes kBlock/2 even sites and the kBlock/2 odd sites of the same checkerboard range (warp-uniform parity), the two
// uses of a matrix are a few CTAs apart and the second one hits in L2.
// Returns the site of launch index i, -1 past the end; the grid is interleaved_blocks(Vh) CTAs of kBlock threads.
__host__ __device__ inline int interleaved_site(int i, int Vh) {
  const int blk = i / kBlock, t = i - blk * kBlock;
  const int par = t >= kBlock / 2 ? 1 : 0;
  const int cb = blk * (kBlock / 2) + t - par * (kBlock / 2);
  return cb < Vh ? par * Vh + cb : -1;
}
inline int interleaved_blocks(int Vh) { return (Vh + kBlock / 2 - 1) / (kBlock / 2); }

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

// Division by a lattice extent: q = (umulhi(m, n) + n) >> l (Granlund-Montgomery), exact for
// 0 <= n < 2^31.  The extents are kernel arguments, so a plain `/` is a 35-instruction reciprocal
// sequence; three of them opened every stencil thread before its first load could issue.
struct FastDiv {
  unsigned m;
  int l;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.l = 0;
  while ((1ll << f.l) < d) f.l++;
  f.m = (unsigned)((((unsigned long long)1 << 32) * (((unsigned long long)1 << f.l) - (unsigned long long)d)) / (unsigned long long)d + 1ull);
  return f;
}
__device__ __forceinline__ int fast_div(int n, const FastDiv f) {
  return (int)((__umulhi(f.m, (unsigned)n) + (unsigned)n) >> f.l);
}

// Local lattice geometry (per GPU).  part[d] != 0 means direction d is split across
// GPUs and neighbours beyond the local extent live in the ghost zone.
struct Geom {
  int L[4];        // local extents
  int Lxh;         // L[0]/2
  int Vh;          // local sites per parity
  int stride;      // colour-vector field stride in sites (Vh padded)
  int gstride;     // ghost-buffer stride in sites (all ghost sites of one half, padded)
  int lstride;     // link field stride in sites (Vh + backward ghost sites, padded)
  int part[4];
  int faceh[4];    // sites per parity in one slice orthogonal to d
  int ghost[4][2]; // Vh + first ghost-buffer site of (d, 0=behind | 1=ahead)
  int lghost[4];   // first backward-ghost site of d in a link half (tail of the link field)
  int origin[4];   // global coordinates of the local origin (even in every direction)
  int G[4];        // global extents
  FastDiv dLxh, dL1, dL2;   // division by Lxh, L[1], L[2]
  int S2, zi;               // Lxh*L[1]; z extent of the interior region (L[2] - 6 when z is partitioned)
  FastDiv dS2, dZi;
};

// Programmatic dependent launch (sm_90+).  A kernel launched with the programmatic-stream-serialisation attribute
// (b200ks.cu launch_k) may be scheduled while its predecessor in the stream is still draining: its CTAs become resident
// as slots free up and stop at pdl_wait() until the predecessor grid has completed and its writes are visible.  What
// that buys the solver loops is the launch latency and the ramp of every kernel boundary (five per CG iteration).
// pdl_launch_dependents() at the top of a kernel lets ITS successor be scheduled as soon as all of this kernel's CTAs
// have started.  Both are no-ops for a kernel launched the ordinary way.  Rule: nothing that reads or writes global
// memory may precede pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Deferred arrival flags of a fused halo push (dslash.cuh push_site_h): the kernel that produced a vector stored its
// boundary sites into the neighbours' ghost buffers with plain (posted) stores; the NEXT kernel on the stream raises
// the neighbours' arrival flags with its first thread.  The kernel boundary in between is what orders the data before
// the flag (a completed grid's writes, peer writes included, are performed system-wide), so the producer needs no
// fence and no ticket -- per-CTA system fences in the producer cost more than the exchange they announced.
struct HaloRaise {
  unsigned long long *flag[4];   // nullptr: direction not partitioned
  unsigned long long seq;        // 0: nothing to raise
};
__device__ __forceinline__ void raise_halo_flags(const HaloRaise &h) {
#if !defined(__CUDA_ARCH__) || __CUDA_ARCH__ >= 600   // (the host-side test builds of the site routines target the default arch)
  if (h.seq == 0) return;
  __threadfence_system();
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (h.flag[k] != nullptr) atomicMax_system(h.flag[k], h.seq);   // flags only ever move forward
#endif
}

struct Coord { int x, y, z, t, xh; };

// cb index -> local coordinates, for a site of the given local parity bit (0 even, 1 odd).
__device__ __forceinline__ Coord site_coord(const Geom &g, int idx, int par) {
  Coord c;
  const int r1 = fast_div(idx, g.dLxh);
  c.xh = idx - r1 * g.Lxh;
  const int r2 = fast_div(r1, g.dL1);
  c.y = r1 - r2 * g.L[1];
  c.t = fast_div(r2, g.dL2);
  c.z = r2 - c.t * g.L[2];
  c.x = 2 * c.xh + ((c.y + c.z + c.t + par) & 1);
  return c;
}

// Index (inside the opposite-parity half) of the neighbour of site (idx,c) displaced by
// h in direction D (h = +-1, +-3).  Non-partitioned directions wrap periodically
// (generic/com_vanilla.c:619-645, ks_spectrum/setup.c:1427-1445).  Partitioned
// directions index the ghost zone.  kLink selects link-field ghosts (backward only).
// kPart = false: the caller knows that no direction is partitioned (single-GPU kernels).
template <int D, bool kLink = false, bool kPart = true>
__device__ __forceinline__ int neighbor(const Geom &g, int idx, const Coord &c, int h) {
  // (h is a compile-time constant at every call site: only the wrap on its side survives)
  if (D == 0) {
    int xn = c.x + h;
    if (h > 0) {
      if (xn >= g.L[0]) xn -= g.L[0];
      if (h > 1 && xn >= g.L[0]) xn -= g.L[0];  // extent 2 with a 3-hop wraps twice
    } else {
      if (xn < 0) xn += g.L[0];
      if (h < -1 && xn < 0) xn += g.L[0];
    }
    return idx - c.xh + (xn >> 1);
  }
  const int coord = (D == 1) ? c.y : (D == 2) ? c.z : c.t;
  const int ext = g.L[D];
  const int sstride = (D == 1) ? g.Lxh : (D == 2) ? g.Lxh * g.L[1] : g.Lxh * g.L[1] * g.L[2];
  int cn = coord + h;
  if (kPart && D >= 2 && g.part[D]) {
    if (cn < 0) {
      // slice -3,-2,-1 -> ghost slice 0,1,2 ; position inside the slice = idx minus this
      // site's own slice offset
      const int inslice = idx - coord * sstride - ((D == 2) ? c.t * sstride * ext : 0);
      const int within = (D == 2) ? inslice + c.t * sstride : inslice;  // (z-slab: keep t-major)
      return (kLink ? g.lghost[D] : g.ghost[D][0]) + (cn + 3) * g.faceh[D] + within;
    }
    if (cn >= ext) {
      const int inslice = idx - coord * sstride - ((D == 2) ? c.t * sstride * ext : 0);
      const int within = (D == 2) ? inslice + c.t * sstride : inslice;
      return g.ghost[D][1] + (cn - ext) * g.faceh[D] + within;
    }
    return idx + h * sstride;
  }
  if (h > 0) {
    if (cn >= ext) cn -= ext;
    if (h > 1 && cn >= ext) cn -= ext;
  } else {
    if (cn < 0) cn += ext;
    if (h < -1 && cn < 0) cn += ext;
  }
  return idx + (cn - coord) * sstride;
}

// ---- complex helpers on T2 -----------------------------------------------------------
template <typename T2, typename T>
__device__ __forceinline__ void cmad(T &re, T &im, const T2 a, const T2 b) {  // += a*b
  re = fma(a.x, b.x, re);
  re = fma(-a.y, b.y, re);
  im = fma(a.x, b.y, im);
  im = fma(a.y, b.x, im);
}
template <typename T2, typename T>
__device__ __forceinline__ void cmad_conj(T &re, T &im, const T2 a, const T2 b) {  // += conj(a)*b
  re = fma(a.x, b.x, re);
  re = fma(a.y, b.y, re);
  im = fma(a.x, b.y, im);
  im = fma(-a.y, b.x, im);
}

// row3 = f * conj(row1 x row2) for a matrix that is (real scalar) x U(3); see load_long.
template <typename T, typename T2>
__device__ __forceinline__ void reconstruct_row3(T2 (&U)[9], const T2 f) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    // c = conj(U[k1] * U[3+k2] - U[k2] * U[3+k1])
    T cr = U[k1].x * U[3 + k2].x;
    cr = fma(-U[k1].y, U[3 + k2].y, cr);
    cr = fma(-U[k2].x, U[3 + k1].x, cr);
    cr = fma(U[k2].y, U[3 + k1].y, cr);
    T ci = U[k1].x * U[3 + k2].y;
    ci = fma(U[k1].y, U[3 + k2].x, ci);
    ci = fma(-U[k2].x, U[3 + k1].y, ci);
    ci = fma(-U[k2].y, U[3 + k1].x, ci);
    ci = -ci;
    U[6 + k].x = fma(f.x, cr, -f.y * ci);
    U[6 + k].y = fma(f.x, ci, f.y * cr);
  }
}

// ---- cache-hinted loads ----------------------------------------------------------------
// Links are streamed once per dslash: evict-first so they do not push the neighbour
// spinors (each re-read 16x) out of L2.  Spinors take the default (read-only) path.
__device__ __forceinline__ double2 ld_stream(const double2 *p) { return __ldcs(p); }
__device__ __forceinline__ float2 ld_stream(const float2 *p) { return __ldcs(p); }
__device__ __forceinline__ double2 ld_keep(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ float2 ld_keep(const float2 *p) { return __ldg(p); }

// ---- deterministic grid reduction ------------------------------------------------------
// Each CTA reduces N doubles with warp shuffles + shared memory and writes one partial;
// the CTA that takes the last ticket sums all partials in a fixed order and stores the
// result.  The order depends only on (gridDim, blockDim), so results are reproducible
// run to run (MILC's site-loop sums are reproducible too; atomics would not be).
struct ReduceWs {
  double *partials;        // [maxBlocks * N]
  unsigned int *counter;   // zero-initialised; reset by the last CTA
};

// Returns true in the CTA that wrote the totals (all of its threads), false elsewhere.
template <int N>
__device__ __forceinline__ bool grid_reduce(double (&v)[N], const ReduceWs ws, double *out) {
  __shared__ double sm[N][kBlock / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      double s = 0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) s += sm[k][w];
      ws.partials[(size_t)blockIdx.x * N + k] = s;
    }
    __threadfence();
    const unsigned ticket = atomicAdd(ws.counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // partial sums of CTAs t, t+128, ...: four CTAs' loads are issued before their (in-order)
  // additions, so the ~gridDim/128 dependent L2 round trips become ~gridDim/512
  double tot_k[N];
#pragma unroll
  for (int k = 0; k < N; k++) tot_k[k] = 0;
  {
    const int G = (int)gridDim.x;
    int b = threadIdx.x;
    for (; b + 3 * kBlock < G; b += 4 * kBlock) {
      double t[4][N];
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < N; k++) t[j][k] = __ldcg(&ws.partials[(size_t)(b + j * kBlock) * N + k]);
#pragma unroll
      for (int j = 0; j < 4; j++)
#pragma unroll
        for (int k = 0; k < N; k++) tot_k[k] += t[j][k];
    }
    for (; b < G; b += kBlock)
#pragma unroll
      for (int k = 0; k < N; k++) tot_k[k] += __ldcg(&ws.partials[(size_t)b * N + k]);
  }
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = tot_k[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) sm[k][warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; w++) tot += sm[k][w];
      out[k] = tot;
    }
  }
  if (threadIdx.x == 0) *ws.counter = 0;
  return true;
}

// ---- two-stage variant (single-GPU solver loops) ----------------------------------------------
// grid_reduce's last CTA adds up gridDim.x x N partial sums with 128 threads: 64 dependent rounds
// of L2 loads at 32^3x64, 10-20 us during which the GPU is otherwise idle (the 16-bit stencil with
// fused dots ran 13 % slower than its bytes explain, the update kernels 40 %).  Inside the solver
// loops the CTAs therefore only store their partial sums; a one-CTA, 1024-thread kernel
// (reduce_finish_kernel, blas.cuh) adds them up in one round and runs the scalar recurrence.
template <int N>
__device__ __forceinline__ void block_partials(double (&v)[N], double *partials) {
  __shared__ double sm[N][kBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sm[k][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kBlock / 32; w++) s += sm[threadIdx.x][w];
    partials[(size_t)blockIdx.x * N + threadIdx.x] = s;
  }
}

}  // namespace b200ks
