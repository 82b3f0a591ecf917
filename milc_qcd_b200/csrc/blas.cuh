// blas.cuh -- fused CG vector updates, device-side scalar recurrences and re-layout kernels.
//
// The solvers never bring a scalar back to the host inside the iteration: every dot
// product lands in a device slot (grid_reduce, common.cuh), the kernels that need
// a = -rsq/pkp etc. recompute it from those slots, and a one-thread "scalar" kernel
// advances the recurrence and raises a stop flag that turns every later kernel of the
// already-enqueued batch into a no-op.  The host polls the flag once per batch.
#pragma once
#include "common.cuh"
#include "comm.cuh"
#include "layout.cuh"

namespace b200ks {

// ---- device-resident solver state ------------------------------------------------------
struct CgState {
  // single-mass CG (generic_ks/d_congrad5_fn_milc.c)
  // red[3] and upd[2] are contiguous on purpose: multi-GPU solves all-reduce them together
  double red[3];       // pkp, c_tr, c_tt of the current iteration (dslash epilogue)
  double upd[2];       // {sum |r|^2 (actual_rsq, :283,318-336), sum |r_s|^2/|x_s|^2} of the
                       // PREVIOUS update = this iteration's oldrsq
  double red_ext[3];   // multi-GPU: the exterior pass's share of red[]
  double upd_next[2];  // reduction target of the update kernel; becomes upd[] afterwards
  double rsq;          // recursive |r|^2 (FEWSUMS expansion value, :339)
  double source_norm;
  double rsqmin, relrsqmin;
  double size_r, size_relr;
  double half_volume;  // global sites per parity, for the relative residue
  int iter;            // iterations done (counts multiplications by M^+M, :223,310)
  int niter;           // restart interval
  int stop;            // 0 run, 1 stop requested (this iteration completes), 2 stopped
  int cur;             // (unused)
  // multi-shift CG (generic_ks/ks_multicg_offset.c)
  int n, n_now, j_low;
  int max_iter;
  double rsq_new, oldrsq, rsqstop;
  double shifts[kMaxShifts];
  double zeta_i[kMaxShifts], zeta_im1[kMaxShifts], zeta_ip1[kMaxShifts];
  double beta_i[kMaxShifts], beta_im1[kMaxShifts], alpha[kMaxShifts];
  // mixed precision (reliable updates)
  double maxrr;        // max recursive |r|^2 since the last reliable update
  double delta2;       // reliable-update threshold squared
  int reliable;        // 1 => host must perform a reliable update
  int pad_;
  const double2 *xrel; // mixed solvers with the Fermilab relative residue: the double part of the solution
                       // (x = xrel + x_lo enters sum |r_s|^2/|x_s|^2); nullptr otherwise
  // multi-shift: per-shift freeze.  The residual of shift j is zeta_j * r, so once
  // zeta_j^2 |r|^2 <= freeze * rsqstop the shift is finished and its two vectors drop out of
  // the update sweep (freeze = 0: never -- the reference iterates every shift to the end,
  // ks_multicg_offset_gpu.c:138-152 "iterate until breakdown").
  double freeze;
  int frozen[kMaxShifts];
};

// ---- plain reductions --------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBlock) norm2_kernel(const typename Vec2<T>::type *v, int stride,
                                                       int n, ReduceWs ws, double *out) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[1] = {0};
  if (i < n) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const auto a = v[(size_t)c * stride + i];
      s[0] += (double)a.x * a.x + (double)a.y * a.y;
    }
  }
  grid_reduce<1>(s, ws, out);
}

template <typename T>
__global__ void __launch_bounds__(kBlock) zero_kernel(typename Vec2<T>::type *v, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  typename Vec2<T>::type z;
  z.x = 0;
  z.y = 0;
#pragma unroll
  for (int c = 0; c < 3; c++) v[(size_t)c * stride + i] = z;
}

template <typename TD, typename TS>
__global__ void __launch_bounds__(kBlock) convert_kernel(typename Vec2<TD>::type *d,
                                                         const typename Vec2<TS>::type *s, int stride,
                                                         int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const auto a = s[(size_t)c * stride + i];
    typename Vec2<TD>::type o;
    o.x = (TD)a.x;
    o.y = (TD)a.y;
    d[(size_t)c * stride + i] = o;
  }
}

// out = a*x + b*y on one parity half (out may alias x or y); y == nullptr: out = a*x
template <typename T>
__global__ void __launch_bounds__(kBlock)
axpby_kernel(typename Vec2<T>::type *out, T a, const typename Vec2<T>::type *x, T b, const typename Vec2<T>::type *y,
             int stride, int n) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const size_t o = (size_t)c * stride + i;
    const T2 xv = x[o];
    T2 r;
    r.x = a * xv.x;
    r.y = a * xv.y;
    if (y != nullptr) {
      const T2 yv = y[o];
      r.x = fma(b, yv.x, r.x);
      r.y = fma(b, yv.y, r.y);
    }
    out[o] = r;
  }
}

// ---- single-mass CG ------------------------------------------------------------------------
// (re)start: ttt already holds D D x - 4m^2 x (= -A x).  r = b + ttt ; p = r ;
// red: |r|^2 and (optionally) sum |r_s|^2/|x_s|^2.       d_congrad5_fn_milc.c:199-218
template <typename T, bool kRel>
__global__ void __launch_bounds__(kBlock)
cg_restart_kernel(const typename Vec2<T>::type *b, const typename Vec2<T>::type *ttt,
                  const typename Vec2<T>::type *x, typename Vec2<T>::type *r,
                  typename Vec2<T>::type *p, int stride, int n, ReduceWs ws, double *out) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[2] = {0, 0};
  if (i < n) {
    double rn = 0, xn = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const T2 bv = b[(size_t)c * stride + i], tv = ttt[(size_t)c * stride + i];
      T2 rv;
      rv.x = bv.x + tv.x;
      rv.y = bv.y + tv.y;
      r[(size_t)c * stride + i] = rv;
      p[(size_t)c * stride + i] = rv;
      rn += (double)rv.x * rv.x + (double)rv.y * rv.y;
      if (kRel) {
        const T2 xv = x[(size_t)c * stride + i];
        xn += (double)xv.x * xv.x + (double)xv.y * xv.y;
      }
    }
    s[0] = rn;
    if (kRel) s[1] = (xn == 0) ? 1.0 : rn / xn;
  }
  grid_reduce<2>(s, ws, out);
}

// One thread: advance the recurrence after the vector update and decide whether the host has
// to look (restart interval reached, or recursive residual under the target).
// Mirrors d_congrad5_fn_milc.c:177-179,310,339,350-354.
__device__ __forceinline__ void cg_scalar_step(CgState *st, int use_rel, int single) {
  if (st->stop) { st->stop = 2; return; }
  const double rsq = st->rsq, oldrsq = st->upd[0];
  const double a = single ? (double)(float)(-rsq / st->red[0]) : -rsq / st->red[0];
  const double rsq_new = oldrsq + 2.0 * a * st->red[1] + a * a * st->red[2];
  // upd_next holds this rank's share; with several GPUs it is summed over ranks together
  // with the next iteration's red[] (or right away when the relative residual is in use)
  st->upd[0] = st->upd_next[0];
  st->upd[1] = st->upd_next[1];
  st->rsq = rsq_new;
  st->iter += 1;
  st->size_r = rsq_new / st->source_norm;
  if (use_rel) st->size_relr = sqrt(st->upd_next[1] / st->half_volume);
  const bool hit_r = (st->rsqmin <= 0 || st->rsqmin > st->size_r);
  const bool hit_rel = (st->relrsqmin <= 0 || st->relrsqmin > st->size_relr);
  if ((st->iter % st->niter == 0) || (hit_r && hit_rel)) st->stop = 1;
  // reliable-update trigger for the mixed-precision solver
  if (st->delta2 > 0) {
    if (rsq_new > st->maxrr) st->maxrr = rsq_new;
    if (rsq_new < st->delta2 * st->maxrr) { st->reliable = 1; st->stop = 1; }
  }
}
__global__ void cg_scalar_kernel(CgState *st, int use_rel, int single) { cg_scalar_step(st, use_rel, single); }

// Second stage of the two-stage reductions (common.cuh block_partials): CTA j adds up slot j's
// per-CTA partial sums and, for the update kernel's sums, advances the scalar recurrence.
// 1024 threads so that the 8192 x 3 partials of a 32^3x64 stencil are one round of loads per
// thread (the same sum by the last CTA of the producing kernel -- 128 threads, 64 dependent
// rounds -- cost 10-20 us per launch, which is what the fused kernels lost against their bytes).
// Fixed order (thread t: CTAs t, t+1024, ...; warp tree; then the 32 warps) => reproducible, and
// the same bits for a block solve and for single solves.  A stopped solver's launch is a no-op.
struct FinishSlot {
  const double *partials;  // value k of CTA b at partials[b * stride + k]
  int stride, nval;        // nval <= 3
  double *out;
  CgState *st;             // scalar step target (scalar_flags != 0)
  const int *stop;
  double *extra;           // partitioned contexts: nextra more values (this rank's share, e.g. the previous
  int nextra;              // update's |r|^2) that ride along in the all-reduce and are replaced by their sums
};
struct FinishArg {
  FinishSlot s[kMaxRhs];
  int nblk;
  int scalar_flags;        // 0: sums only; else bit 0 on, bit 1 use_rel, bit 2 single (cg_scalar_step)
  HaloRaise raise;         // arrival flags of the halo the kernel before this one pushed (common.cuh)
};

// ---- flag-based all-reduce over the ranks of a partitioned context (comm.cuh RedBox) -----------
// Block-wide: v[0..n) in shared memory holds this rank's values on entry and the sums (kMax: the
// maxima) over all ranks on return, bit-identical on every rank.  n <= 8, blockDim >= 8 * nranks.
template <bool kMax>
__device__ __forceinline__ void p2p_allreduce_block(double *v, int n, const RedComm &rc) {
  RedBox *own = rc.box[rc.rank];
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = own->count + 1;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int buf = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < rc.nranks * 8) {
    const int q = t >> 3, k = t & 7;
    if (k < n) {
      volatile double *dst = &rc.box[q]->val[buf][rc.rank][k];
      *dst = v[k];
    }
  }
  __threadfence_system();   // every thread's mailbox stores before ...
  __syncthreads();          // ... the flag stores below (other threads')
  if (t < rc.nranks) {
    unsigned long long *f = &rc.box[t]->flag[rc.rank];
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(seq) : "memory");
    const unsigned long long *mine = &own->flag[t];
    const long long t0 = clock64();
    for (;;) {
      unsigned long long x;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(x) : "l"(mine) : "memory");
      if (x >= seq) break;
      if (clock64() - t0 > rc.timeout) {
        *rc.err = 100 + t;
        break;
      }
    }
  }
  __syncthreads();
  if (t < n) {
    double s = kMax ? -1.7976931348623157e308 : 0.0;
    for (int q = 0; q < rc.nranks; q++) {
      const volatile double *src = &own->val[buf][q][t];
      const double x = *src;
      s = kMax ? fmax(s, x) : s + x;
    }
    v[t] = s;
  }
  if (t == 0) own->count = seq;
  __syncthreads();
}

// stand-alone form: d[0..n) <- sum (max) over ranks
constexpr int kAllreduceThreads = 128;
template <bool kMax>
__global__ void __launch_bounds__(kAllreduceThreads) p2p_allreduce_kernel(double *d, int n, const RedComm rc, const int *stop) {
  if (stop != nullptr && *stop) return;
  __shared__ double v[8];
  if (threadIdx.x < 8) v[threadIdx.x] = threadIdx.x < n ? d[threadIdx.x] : 0.0;
  __syncthreads();
  p2p_allreduce_block<kMax>(v, n, rc);
  if (threadIdx.x < n) d[threadIdx.x] = v[threadIdx.x];
}

constexpr int kFinishThreads = 1024;
// kComm: the context is partitioned; slot 0's sums (and its `extra` values) are all-reduced over the
// ranks before they are stored (one slot per launch).
template <bool kComm>
__global__ void __launch_bounds__(kFinishThreads) reduce_finish_kernel(const FinishArg a, const RedComm rc) {
  pdl_launch_dependents();
  pdl_wait();
  const FinishSlot &f = a.s[blockIdx.x];
  if (f.stop != nullptr && *f.stop) return;
  if (blockIdx.x == 0 && threadIdx.x == 0) raise_halo_flags(a.raise);
  __shared__ double sm[3][kFinishThreads / 32];
  __shared__ double tot[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s[3] = {0, 0, 0};
  for (int b0 = threadIdx.x; b0 < a.nblk; b0 += 4 * kFinishThreads) {
    double t[4][3];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int b = b0 + j * kFinishThreads;
#pragma unroll
      for (int k = 0; k < 3; k++) t[j][k] = (b < a.nblk && k < f.nval) ? __ldcg(&f.partials[(size_t)b * f.stride + k]) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int k = 0; k < 3; k++) s[k] += t[j][k];
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) sm[k][warp] = s[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      double v = sm[k][lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) tot[k] = v;
    }
  }
  if (kComm) {
    if (threadIdx.x < f.nextra) tot[f.nval + threadIdx.x] = f.extra[threadIdx.x];
    __syncthreads();
    p2p_allreduce_block<false>(tot, f.nval + f.nextra, rc);
    if (threadIdx.x < f.nextra) f.extra[threadIdx.x] = tot[f.nval + threadIdx.x];
  } else {
    __syncthreads();
  }
  if (threadIdx.x < f.nval) f.out[threadIdx.x] = tot[threadIdx.x];
  if (a.scalar_flags) {
    __syncthreads();   // the scalar step reads what the threads above have just stored
    if (threadIdx.x == 0) cg_scalar_step(f.st, (a.scalar_flags >> 1) & 1, (a.scalar_flags >> 2) & 1);
  }
}

// One fused update per iteration (the reference's FEWSUMS arithmetic, :301-345,363-367):
//   a = -rsq/pkp ; rsq' = oldrsq + 2a c_tr + a^2 c_tt ; b = rsq'/oldrsq
//   x += a p ; r += a ttt ; p = r + b p ; actual' = sum |r|^2 (summed for the NEXT iteration)
// Every thread derives a, b from the device slots; nothing is mutated here except the
// reduction target actual[next].
template <typename T, bool kRel>
__global__ void __launch_bounds__(kBlock)
cg_update_kernel(typename Vec2<T>::type *x, typename Vec2<T>::type *r, typename Vec2<T>::type *p,
                 const typename Vec2<T>::type *ttt, int stride, int n, CgState *st, ReduceWs ws, int fuse_scalar) {
  using T2 = typename Vec2<T>::type;
  pdl_launch_dependents();
  pdl_wait();
  if (st->stop) return;
  const double rsq = st->rsq, oldrsq = st->upd[0];
  const double pkp = st->red[0], c_tr = st->red[1], c_tt = st->red[2];
  const T a = (T)(-rsq / pkp);
  const double rsq_new = oldrsq + 2.0 * (double)a * c_tr + (double)a * (double)a * c_tt;
  const T bb = (T)(rsq_new / oldrsq);
  const int i = blockIdx.x * kBlock + threadIdx.x;
  const double2 *xrel = kRel ? st->xrel : nullptr;
  double s[2] = {0, 0};
  if (i < n) {
    T rn = 0, xn = 0;   // per-site sums in the working precision, summed over sites in double
    double xn2 = 0;     // mixed solvers: |x_double + x_lo|^2
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      T2 xv = x[o], rv = r[o], pv = p[o];
      const T2 tv = ttt[o];
      xv.x = fma(a, pv.x, xv.x);
      xv.y = fma(a, pv.y, xv.y);
      rv.x = fma(a, tv.x, rv.x);
      rv.y = fma(a, tv.y, rv.y);
      pv.x = fma(bb, pv.x, rv.x);
      pv.y = fma(bb, pv.y, rv.y);
      x[o] = xv;
      r[o] = rv;
      p[o] = pv;
      rn = fma(rv.x, rv.x, fma(rv.y, rv.y, rn));
      if (kRel) {
        if (xrel != nullptr) {
          const double2 xd = xrel[o];
          const double tx = xd.x + (double)xv.x, ty = xd.y + (double)xv.y;
          xn2 += tx * tx + ty * ty;
        } else {
          xn = fma(xv.x, xv.x, fma(xv.y, xv.y, xn));
        }
      }
    }
    s[0] = rn;
    if (kRel) {
      if (xrel != nullptr) s[1] = (xn2 == 0) ? 1.0 : (double)rn / xn2;
      else s[1] = (xn == 0) ? 1.0 : (double)rn / (double)xn;
    }
  }
  // safe although other CTAs read st->upd at their start: the last ticket is taken only
  // after every CTA has passed that read.  For the same reason the CTA that writes the totals
  // can advance the scalar recurrence right away (fuse_scalar: bit 0 on, bit 1 use_rel, bit 2
  // single) instead of leaving it to a one-thread kernel -- one launch less per iteration.
  if (fuse_scalar & 8) {   // two-stage: reduce_finish_kernel sums the partials and advances the recurrence
    block_partials<2>(s, ws.partials);
    return;
  }
  const bool last = grid_reduce<2>(s, ws, st->upd_next);
  if (last && fuse_scalar && threadIdx.x == 0) cg_scalar_step(st, (fuse_scalar >> 1) & 1, (fuse_scalar >> 2) & 1);
}

// One thread: advance the recurrence after cg_update_kernel and decide whether the host
// has to look (restart interval reached, or recursive residual under the target).
// Mirrors d_congrad5_fn_milc.c:177-179,310,339,350-354.
// multi-GPU: fold the exterior pass's partial sums into red[] before the all-reduce
__global__ void combine_red_kernel(CgState *st, int n) {
  if (st->stop) return;
  for (int k = 0; k < n; k++) st->red[k] += st->red_ext[k];
}


// ---- mixed precision: double outer solution, single inner Krylov vectors -----------------------
// x(double) += x_lo(float) ; x_lo = 0      (before every true-residual evaluation)
__global__ void __launch_bounds__(kBlock)
mixed_accumulate_kernel(double2 *x, float2 *x_lo, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const size_t o = (size_t)c * stride + i;
    double2 xv = x[o];
    const float2 lv = x_lo[o];
    xv.x += (double)lv.x;
    xv.y += (double)lv.y;
    x[o] = xv;
    x_lo[o] = make_float2(0.f, 0.f);
  }
}

// Reliable update: ttt holds D D x - 4m^2 x in double.  r_true = b + ttt replaces the
// single-precision recursive residual; the search direction keeps its Krylov history and is
// only shifted by the residual correction (p = r + b p_old with the corrected r).
// first != 0: start of the solve, p = r.
__global__ void __launch_bounds__(kBlock)
mixed_reliable_kernel(const double2 *b, const double2 *ttt, float2 *r_lo, float2 *p_lo, int stride, int n,
                      int first, ReduceWs ws, double *out, const double2 *xrel) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[2] = {0, 0};
  if (i < n) {
    double num = 0, den = 0;   // xrel != nullptr: the Fermilab relative residue of the TRUE residual (d_congrad5_fn_milc.c:37-56)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      const double2 bv = b[o], tv = ttt[o];
      const double rx = bv.x + tv.x, ry = bv.y + tv.y;
      const float2 ro = r_lo[o];
      const float2 rn = make_float2((float)rx, (float)ry);
      r_lo[o] = rn;
      if (first) {
        p_lo[o] = rn;
      } else {
        float2 pv = p_lo[o];
        pv.x += rn.x - ro.x;
        pv.y += rn.y - ro.y;
        p_lo[o] = pv;
      }
      s[0] += rx * rx + ry * ry;
      if (xrel != nullptr) {
        const double2 xv = xrel[o];
        num += rx * rx + ry * ry;
        den += xv.x * xv.x + xv.y * xv.y;
      }
    }
    if (xrel != nullptr) s[1] = (den == 0) ? 1.0 : num / den;
  }
  grid_reduce<2>(s, ws, out);
}

// ---- multi-shift CG --------------------------------------------------------------------------
// r += beta_low * ttt ; rsq_new = |r|^2          ks_multicg_offset.c:365-368
template <typename T>
__global__ void __launch_bounds__(kBlock)
ms_resid_kernel(typename Vec2<T>::type *r, const typename Vec2<T>::type *ttt, int stride, int n,
                CgState *st, ReduceWs ws) {
  using T2 = typename Vec2<T>::type;
  if (st->stop) return;
  const T beta = (T)(-st->rsq / st->red[0]);
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[1] = {0};
  if (i < n) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      T2 rv = r[o];
      const T2 tv = ttt[o];
      rv.x = fma(beta, tv.x, rv.x);
      rv.y = fma(beta, tv.y, rv.y);
      r[o] = rv;
      s[0] += (double)rv.x * rv.x + (double)rv.y * rv.y;
    }
  }
  grid_reduce<1>(s, ws, &st->rsq_new);
}

// One thread: the zeta/beta/alpha recurrences with the reference's zero guards and
// trailing-shift dropping, the convergence test and the scroll.
// ks_multicg_offset.c:322-355,381,429-444,456-460.
__global__ void ms_scalar_kernel(CgState *st) {
  if (st->stop) { st->stop = 2; return; }
  const int jl = st->j_low;
  const double rsq = st->rsq, pkp = st->red[0];
  st->oldrsq = rsq;
  st->iter += 1;
  st->beta_i[jl] = -rsq / pkp;
  st->zeta_ip1[jl] = 1.0;
  for (int j = 0; j < st->n_now; j++) {
    if (j == jl) continue;
    st->zeta_ip1[j] = st->zeta_i[j] * st->zeta_im1[j] * st->beta_im1[jl];
    const double c1 = st->beta_i[jl] * st->alpha[jl] * (st->zeta_im1[j] - st->zeta_i[j]);
    const double c2 = st->zeta_im1[j] * st->beta_im1[jl] * (1.0 + st->shifts[j] * st->beta_i[jl]);
    if (c1 + c2 != 0.0) st->zeta_ip1[j] /= c1 + c2;
    else st->zeta_ip1[j] = 0.0;
    if (st->zeta_i[j] != 0.0) {
      st->beta_i[j] = st->beta_i[jl] * st->zeta_ip1[j] / st->zeta_i[j];
    } else {
      st->zeta_ip1[j] = 0.0;
      st->beta_i[j] = 0.0;
      if (j == st->n_now - 1 && j > jl) st->n_now--;
    }
  }
  const double rsq_new = st->rsq_new;
  st->rsq = rsq_new;
  st->size_r = rsq_new / st->source_norm;
  if (st->rsqstop > 0 && rsq_new <= st->rsqstop) { st->stop = 1; return; }
  if (st->freeze > 0)
    for (int j = 0; j < st->n_now; j++)
      if (j != jl && !st->frozen[j] && st->zeta_ip1[j] * st->zeta_ip1[j] * rsq_new <= st->freeze * st->rsqstop) st->frozen[j] = 2;
  st->alpha[jl] = rsq_new / rsq;
  for (int j = 0; j < st->n_now; j++) {
    if (j == jl) continue;
    if (st->zeta_i[j] * st->beta_i[jl] != 0.0)
      st->alpha[j] = st->alpha[jl] * st->zeta_ip1[j] * st->beta_i[j] / (st->zeta_i[j] * st->beta_i[jl]);
    else st->alpha[j] = 0.0;
  }
  if (st->iter >= st->max_iter) st->stop = 1;
}

// Scroll after the vector update has consumed zeta_ip1/alpha (kept separate so the vector
// kernel reads a consistent set).
__global__ void ms_scroll_kernel(CgState *st) {
  if (st->stop) return;
  for (int j = 0; j < st->n_now; j++)
    if (st->frozen[j] == 2) st->frozen[j] = 1;   // its last x update has just been applied
  for (int j = 0; j < st->n_now; j++) {
    st->beta_im1[j] = st->beta_i[j];
    st->zeta_im1[j] = st->zeta_i[j];
    st->zeta_i[j] = st->zeta_ip1[j];
  }
}

struct MsPtrs {
  void *x[kMaxShifts];
  void *pm[kMaxShifts];
};

// All shifts in one pass:  x_j += beta_j pm_j ;  pm_j = zeta_ip1_j r + alpha_j pm_j.
// (ks_multicg_offset.c:358-362,446-453, fused: each x_j, pm_j is read and written once
// per iteration, r is read once for all shifts.)  When stop == 1 (converged in this
// iteration) only the x update is applied, which is the state the reference returns.
template <typename T>
__global__ void __launch_bounds__(kBlock)
ms_update_kernel(const MsPtrs ptrs, const typename Vec2<T>::type *r, int stride, int n,
                 const CgState *st) {
  using T2 = typename Vec2<T>::type;
  const int stop = st->stop;
  if (stop == 2) return;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  T2 rv[3];
#pragma unroll
  for (int c = 0; c < 3; c++) rv[c] = r[(size_t)c * stride + i];
  const int nn = st->n_now;
  for (int j = 0; j < nn; j++) {
    if (st->frozen[j] == 1) continue;   // finished shift: x_j is final, pm_j is dead
    const T beta = (T)st->beta_i[j], zeta = (T)st->zeta_ip1[j], alpha = (T)st->alpha[j];
    T2 *x = (T2 *)ptrs.x[j];
    T2 *pm = (T2 *)ptrs.pm[j];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      T2 xv = x[o], pv = pm[o];
      xv.x = fma(beta, pv.x, xv.x);
      xv.y = fma(beta, pv.y, xv.y);
      x[o] = xv;
      if (stop == 0) {
        pv.x = fma(alpha, pv.x, zeta * rv[c].x);
        pv.y = fma(alpha, pv.y, zeta * rv[c].y);
        pm[o] = pv;
      }
    }
  }
}

// r = b ; pm_j = b ; x_j = 0 ; source_norm               ks_multicg_offset.c:223-235
template <typename T>
__global__ void __launch_bounds__(kBlock)
ms_init_kernel(const MsPtrs ptrs, int nshift, const typename Vec2<T>::type *b,
               typename Vec2<T>::type *r, int stride, int n, ReduceWs ws, double *out) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double s[1] = {0};
  if (i < n) {
    T2 z;
    z.x = 0;
    z.y = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const size_t o = (size_t)c * stride + i;
      const T2 bv = b[o];
      r[o] = bv;
      s[0] += (double)bv.x * bv.x + (double)bv.y * bv.y;
      for (int j = 0; j < nshift; j++) {
        ((T2 *)ptrs.x[j])[o] = z;
        ((T2 *)ptrs.pm[j])[o] = bv;
      }
    }
  }
  grid_reduce<1>(s, ws, out);
}

// ---- MILC host layout <-> device layout ---------------------------------------------------------
// Colour vectors: host su3_vector[V] (AoS, 6 reals/site) for one parity block starting at
// host site offset `hoff`  ->  device SoA.  TH = host real type, T = device real type.
template <typename T, typename TH>
__global__ void __launch_bounds__(kBlock)
pack_vec_kernel(typename Vec2<T>::type *d, const TH *h, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const TH *s = h + (size_t)6 * i;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    typename Vec2<T>::type o;
    o.x = (T)s[2 * c];
    o.y = (T)s[2 * c + 1];
    d[(size_t)c * stride + i] = o;
  }
}
template <typename T, typename TH>
__global__ void __launch_bounds__(kBlock)
unpack_vec_kernel(TH *h, const typename Vec2<T>::type *d, int stride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  TH *s = h + (size_t)6 * i;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const auto o = d[(size_t)c * stride + i];
    s[2 * c] = (TH)o.x;
    s[2 * c + 1] = (TH)o.y;
  }
}
// ncomp = 4 * (complex numbers per link): 36 for full matrices, 28 for compressed long links
template <typename TD, typename TS>
__global__ void __launch_bounds__(kBlock)
convert_link_kernel(typename Vec2<TD>::type *d, const typename Vec2<TS>::type *s, int lstride, int n, int ncomp) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
#pragma unroll 4
  for (int m = 0; m < ncomp; m++) {
    const auto a = s[(size_t)m * lstride + i];
    typename Vec2<TD>::type o;
    o.x = (TD)a.x;
    o.y = (TD)a.y;
    d[(size_t)m * lstride + i] = o;
  }
}

// ---- long-link compression (two rows + one complex U(3) factor, see dslash.cuh load_long) ------
// Least-squares factor f with row3 = f * conj(row1 x row2), and the relative misfit
// max_k |row3_k - f c_k| / |row1| that decides whether compression is legitimate.
template <typename T>
__device__ __forceinline__ void long_factor(const typename Vec2<T>::type (&U)[9], double &fx, double &fy, double &dev) {
  double cx[3], cy[3], nn = 0, px = 0, py = 0, n1 = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    const double ar = U[k1].x, ai = U[k1].y, br = U[3 + k2].x, bi = U[3 + k2].y;
    const double er = U[k2].x, ei = U[k2].y, gr = U[3 + k1].x, gi = U[3 + k1].y;
    cx[k] = (ar * br - ai * bi) - (er * gr - ei * gi);
    cy[k] = -((ar * bi + ai * br) - (er * gi + ei * gr));
    nn += cx[k] * cx[k] + cy[k] * cy[k];
    // row3_k * conj(c_k)
    px += (double)U[6 + k].x * cx[k] + (double)U[6 + k].y * cy[k];
    py += (double)U[6 + k].y * cx[k] - (double)U[6 + k].x * cy[k];
    n1 += (double)U[k].x * U[k].x + (double)U[k].y * U[k].y;
  }
  if (nn == 0) {  // zero link (padding): any f works
    fx = fy = 0;
    double r3 = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) r3 += (double)U[6 + k].x * U[6 + k].x + (double)U[6 + k].y * U[6 + k].y;
    dev = (r3 == 0) ? 0.0 : 1.0;
    return;
  }
  fx = px / nn;
  fy = py / nn;
  double d2 = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double rx = U[6 + k].x - (fx * cx[k] - fy * cy[k]);
    const double ry = U[6 + k].y - (fx * cy[k] + fy * cx[k]);
    d2 = fmax(d2, rx * rx + ry * ry);
  }
  dev = sqrt(d2 / n1);
}

// max misfit over all links of one parity half -> *out (bits of a non-negative double, atomicMax)
template <typename T>
__global__ void __launch_bounds__(kBlock)
long_deviation_kernel(const typename Vec2<T>::type *lng, int lstride, int n, unsigned long long *out) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  double worst = 0;
  if (i < n) {
    for (int mu = 0; mu < 4; mu++) {
      T2 U[9];
#pragma unroll
      for (int e = 0; e < 9; e++) U[e] = lng[(size_t)(mu * 9 + e) * lstride + i];
      double fx, fy, dev;
      long_factor<T>(U, fx, fy, dev);
      if (!(dev <= worst)) worst = (dev == dev) ? dev : 1.0;   // NaN counts as a failure
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0 && worst > 0) atomicMax(out, (unsigned long long)__double_as_longlong(worst));
}

// [mu][9][site] -> [mu][7][site]
template <typename T>
__global__ void __launch_bounds__(kBlock)
compress_long_kernel(typename Vec2<T>::type *dst, const typename Vec2<T>::type *src, int lstride, int n) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  for (int mu = 0; mu < 4; mu++) {
    T2 U[9];
#pragma unroll
    for (int e = 0; e < 9; e++) U[e] = src[(size_t)(mu * 9 + e) * lstride + i];
    double fx, fy, dev;
    long_factor<T>(U, fx, fy, dev);
#pragma unroll
    for (int e = 0; e < 6; e++) dst[(size_t)(mu * 7 + e) * lstride + i] = U[e];
    T2 f;
    f.x = (T)fx;
    f.y = (T)fy;
    dst[(size_t)(mu * 7 + 6) * lstride + i] = f;
  }
}

// compressed device long links -> host su3_matrix[4*V] layout
template <typename T, typename TH>
__global__ void __launch_bounds__(kBlock)
unpack_long7_kernel(TH *h, const typename Vec2<T>::type *d, int lstride, int n) {
  using T2 = typename Vec2<T>::type;
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  TH *s = h + (size_t)72 * i;
  for (int mu = 0; mu < 4; mu++) {
    T2 U[9];
#pragma unroll
    for (int e = 0; e < 6; e++) U[e] = d[(size_t)(mu * 7 + e) * lstride + i];
    reconstruct_row3<T, T2>(U, d[(size_t)(mu * 7 + 6) * lstride + i]);
#pragma unroll
    for (int e = 0; e < 9; e++) {
      s[2 * (mu * 9 + e)] = (TH)U[e].x;
      s[2 * (mu * 9 + e) + 1] = (TH)U[e].y;
    }
  }
}

}  // namespace b200ks
