// fermion_force.cu -- host side of the HISQ fermion force (force.cuh):
// buffers, uploads, the chain of launches, and the C ABI entry point behind qudaHisqForce
// (generic_ks/fermion_force_hisq_multi.c:2169-2290).
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#include "force.cuh"
#include "internal.h"
#include "layout.cuh"

using namespace b200ks;
using namespace b200ks_host;

namespace {

// Defaults of the A/B switches read in b200ks_hisq_force.  Measured at 32^3 x 64, 9 terms (profiles/force_ab_r02u.json,
// profiles/ncu_fbwd_r02v.txt): uploading V and U beside the W-level chain takes 0.09 s off a 0.42 s call; the two-role
// form of the backward staple pass runs in 847 us against the fused body's 868 (both with the interleaved site order),
// i.e. doubling the resident warps buys 2 % -- the pass is not waiting for occupancy -- so the fused body stays.
constexpr int kDefaultForceForm = 0;         // B200KS_FORCE_SPLIT
constexpr bool kDefaultForceOverlap = true;  // B200KS_FORCE_OVERLAP

// one thread per site: the kernel is the __host__ __device__ site functor of force.cuh plus the launch.
// vh_il != 0 (functors over the 2 Vh sites only): the two parities interleaved CTA by CTA (common.cuh interleaved_site)
template <class F>
__global__ void __launch_bounds__(kBlock) force_site_kernel(const F fn, int n, int vh_il) {
  int i = blockIdx.x * kBlock + threadIdx.x;
  if (vh_il) i = interleaved_site(i, vh_il);
  else if (i >= n) i = -1;
  if (i >= 0) fn(i);
}

// the same with a register cap (kMinBlocks 128-thread CTAs per SM: 4 = 128 registers, 3 = 168) for functors that
// declare a static kMinBlocks member
template <class F>
__global__ void __launch_bounds__(kBlock, F::kMinBlocks) force_site_kernel2(const F fn, int n, int vh_il) {
  int i = blockIdx.x * kBlock + threadIdx.x;
  if (vh_il) i = interleaved_site(i, vh_il);
  else if (i >= n) i = -1;
  if (i >= 0) fn(i);
}
template <class F, class = void>
struct WantsCap : std::false_type {};
template <class F>
struct WantsCap<F, std::void_t<decltype(F::kMinBlocks)>> : std::true_type {};

// Two threads per site (force.cuh StapleBwdPairSite): warps 0-1 of a CTA run role_link for kBlock/2 sites, warps 2-3
// role_u for the same sites; the one matrix that crosses goes through shared memory.  vh_il != 0: the CTA's sites
// are kBlock/4 even sites and the kBlock/4 odd sites of the same checkerboard range (parity warp-uniform).
constexpr int kPairSites = kBlock / 2;
template <class F>
__global__ void __launch_bounds__(kBlock, F::kMinBlocks) force_pair_kernel(const F fn, int n, int vh_il) {
  __shared__ double2 hand[9][kPairSites];   // role_link's sixth contribution
  __shared__ double2 sum[9][kPairSites];    // role_u's running sum
  const int role = threadIdx.x >= kPairSites ? 1 : 0;
  const int t = threadIdx.x - role * kPairSites;
  int z;
  if (vh_il) {
    const int par = t >= kPairSites / 2 ? 1 : 0;
    const int cb = blockIdx.x * (kPairSites / 2) + t - par * (kPairSites / 2);
    z = cb < vh_il ? par * vh_il + cb : -1;
  } else {
    z = blockIdx.x * kPairSites + t;
    if (z >= n) z = -1;
  }
  if (z >= 0) {
    if (role == 0) {
      fn.role_link(z, [&](const force::Mat &t4) {
#pragma unroll
        for (int k = 0; k < 9; k++) hand[k][t] = t4.e[k];
      });
    } else {
      bool first = true;
      fn.role_u(z, [&](const force::Mat &term) {
#pragma unroll
        for (int k = 0; k < 9; k++) {
          double2 v = term.e[k];
          if (!first) {   // (0 + term would turn a -0 into +0: the first term is stored)
            const double2 s = sum[k][t];
            v.x = s.x + v.x;
            v.y = s.y + v.y;
          }
          sum[k][t] = v;
        }
        first = false;
      });
    }
  }
  __syncthreads();
  if (z >= 0 && role == 1) {
    force::Mat gu, t4;
#pragma unroll
    for (int k = 0; k < 9; k++) {
      gu.e[k] = sum[k][t];
      t4.e[k] = hand[k][t];
    }
    fn.finish_u(z, gu, t4);
  }
}

struct DeviceExec {
  b200ks_ctx *c;
  int nsites = 0;   // 2 Vh: launches over exactly the sites may be interleaved
  int order = 0;    // site_order() of this call
  template <class F>
  void run(int n, const F &fn) {
    const int vh_il = (order && n == nsites) ? nsites / 2 : 0;
    const int grid = vh_il ? interleaved_blocks(vh_il) : nblocks(n);
    if constexpr (std::is_same<F, force::StapleBwdPairSite<3>>::value || std::is_same<F, force::StapleBwdPairSite<4>>::value) {
      const int per = vh_il ? kPairSites / 2 : kPairSites;
      const int m = vh_il ? vh_il : n;
      force_pair_kernel<F><<<(m + per - 1) / per, kBlock, 0, stream(c)>>>(fn, n, vh_il);
    } else if constexpr (WantsCap<F>::value) {
      force_site_kernel2<F><<<grid, kBlock, 0, stream(c)>>>(fn, n, vh_il);
    } else {
      force_site_kernel<F><<<grid, kBlock, 0, stream(c)>>>(fn, n, vh_il);
    }
    count_launch(c);
  }
};

template <typename TH>
__global__ void __launch_bounds__(kBlock) widen_kernel(double *d, const TH *s, size_t n) {
  const size_t i = (size_t)blockIdx.x * kBlock + threadIdx.x;
  if (i < n) d[i] = (double)s[i];
}

struct Scratch {   // freed on every exit path
  b200ks_ctx *c;
  void *p = nullptr;
  size_t bytes = 0;
  explicit Scratch(b200ks_ctx *c_) : c(c_) {}
  ~Scratch() {
    if (p) {
      cudaStreamSynchronize(stream(c));
      dev_release(c, p, bytes);
    }
  }
};

// host su3_matrix[4*V] (MILC order) -> full-lattice matrix field (36 planes), stream-ordered on `st`
// (the context's stream, or the upload stream that runs beside it) and complete on return
int field_to_dev(b200ks_ctx *c, cudaStream_t st, double2 *dst, size_t fs, const void *host, int host_prec) {
  const int Vh = geom(c).Vh;
  const size_t hs = host_prec == 2 ? 8 : 4;
  const size_t half_bytes = (size_t)Vh * 72 * hs;
  void *stage = nullptr;
  CHK(stage_get(c, half_bytes, &stage));
  for (int p = 0; p < 2; p++) {
    CHK(h2d_on(c, st, stage, (const char *)host + (size_t)p * half_bytes, half_bytes));
    if (host_prec == 2) pack_link_kernel<double, double><<<nblocks(Vh), kBlock, 0, st>>>(dst + (size_t)p * Vh, (const double *)stage, (int)fs, Vh);
    else pack_link_kernel<double, float><<<nblocks(Vh), kBlock, 0, st>>>(dst + (size_t)p * Vh, (const float *)stage, (int)fs, Vh);
    count_launch(c);
    CU(cudaStreamSynchronize(st));   // the staging buffer is reused
  }
  return check_launch("pack_link_kernel");
}

struct UploadStream {   // destroyed on every exit path
  cudaStream_t s = nullptr;
  cudaEvent_t done = nullptr;
  ~UploadStream() {
    if (s) {
      cudaStreamSynchronize(s);
      cudaStreamDestroy(s);
    }
    if (done) cudaEventDestroy(done);
  }
};

}  // namespace

extern "C" int b200ks_hisq_force(b200ks_ctx *c, int nterms, int num_naik_terms, const double *coeff, const void *const *multi_x,
                                 const double *level2_coeff, const double *fat7_coeff, const void *wlink, const void *vlink,
                                 const void *ulink, double eps, double force_filter, void *momentum, int host_prec) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  if (!c || nterms < 1 || !coeff || !multi_x || !level2_coeff || !fat7_coeff || !wlink || !vlink || !ulink || !momentum)
    return fail(B200KS_EINVAL, "b200ks_hisq_force: null argument");
  if (host_prec != 1 && host_prec != 2) return fail(B200KS_EINVAL, "host_prec must be 1 or 2");
  if (!(force_filter >= 0.0)) return fail(B200KS_EINVAL, "b200ks_hisq_force: negative force filter");
  if (num_naik_terms < 0 || num_naik_terms > nterms) return fail(B200KS_EINVAL, "b200ks_hisq_force: num_naik_terms out of range");
  if (partitioned(c)) return fail(B200KS_ESTATE, "fermion force: single-GPU contexts only");
  CU(cudaSetDevice(device(c)));
  const Geom &g = geom(c);
  const int n = 2 * g.Vh;
  force::ForceBufs b;
  bool overlap = kDefaultForceOverlap;
  {   // A/B switches.  B200KS_FORCE_SPLIT: the backward staple passes as four small kernels (1) or as two roles of one
      // kernel (2: compiled for 128 registers, 3: for 168) instead of the fused body (0); B200KS_FORCE_OVERLAP: V and U travel while the W-level chain runs
    const char *e = getenv("B200KS_FORCE_SPLIT");
    const int form = e ? atoi(e) : kDefaultForceForm;
    b.split = form == 1;
    b.pair = form == 2 ? 4 : form == 3 ? 3 : 0;
    if ((e = getenv("B200KS_FORCE_OVERLAP")) != nullptr) overlap = atoi(e) != 0;
  }
  for (int d = 0; d < 4; d++) b.g.L[d] = g.L[d];
  b.g.Vh = g.Vh;
  b.nsites = n;
  b.fs = ((size_t)n + 63) / 64 * 64;
  const size_t m4 = 36 * b.fs, m1 = 9 * b.fs;
  Scratch mats(c), vec(c), mom(c);
  mats.bytes = (7 * m4 + 4 * m1) * sizeof(double2);
  CHK(dev_alloc(c, &mats.p, mats.bytes));
  vec.bytes = (size_t)n * 6 * sizeof(double);
  CHK(dev_alloc(c, &vec.p, vec.bytes));
  mom.bytes = (size_t)n * 40 * sizeof(double);
  CHK(dev_alloc(c, &mom.p, mom.bytes));
  double2 *p = (double2 *)mats.p;
  b.U = p; p += m4;
  b.V = p; p += m4;
  b.W = p; p += m4;
  b.gfat = p; p += m4;
  b.glng = p; p += m4;
  b.gW = p; p += m4;
  b.gU = p; p += m4;
  b.st3 = p; p += m1;
  b.st5 = p; p += m1;
  b.g3 = p; p += m1;
  b.g5 = p; p += m1;
  // W first: the outer products need no links and the level-2 chain needs W only
  if (!overlap) {
    CHK(field_to_dev(c, stream(c), b.U, b.fs, ulink, host_prec));
    CHK(field_to_dev(c, stream(c), b.V, b.fs, vlink, host_prec));
  }
  CHK(field_to_dev(c, stream(c), b.W, b.fs, wlink, host_prec));
  DeviceExec x{c, n, site_order()};
  x.run(n, force::ZeroSite{b.gfat, b.fs, 36});
  x.run(n, force::ZeroSite{b.glng, b.fs, 36});
  if (num_naik_terms > 0) {   // the Naik-epsilon terms' outer products go straight to the W level (force_chain)
    x.run(n, force::ZeroSite{b.gW, b.fs, 36});
    x.run(n, force::ZeroSite{b.gU, b.fs, 36});
  }
  // outer products, one term at a time: colour vectors stay in MILC's host order (6 reals per site)
  for (int j = 0; j < nterms; j++) {
    if (!multi_x[j]) return fail(B200KS_EINVAL, "b200ks_hisq_force: null vector");
    if (host_prec == 2) {
      CHK(h2d(c, vec.p, multi_x[j], vec.bytes));
    } else {
      void *stage = nullptr;
      CHK(stage_get(c, (size_t)n * 6 * sizeof(float), &stage));
      CHK(h2d(c, stage, multi_x[j], (size_t)n * 6 * sizeof(float)));
      widen_kernel<float><<<nblocks(6 * n), kBlock, 0, stream(c)>>>((double *)vec.p, (const float *)stage, (size_t)6 * n);
      count_launch(c);
    }
    x.run(n, force::OprodSite{b.g, b.gfat, b.glng, b.fs, (const double *)vec.p, coeff[2 * j], coeff[2 * j + 1]});
    const int i = j - (nterms - num_naik_terms);   // the last num_naik_terms fields, weights coeff[nterms + i]
    if (i >= 0)
      x.run(n, force::OprodSite{b.g, b.gW, b.gU, b.fs, (const double *)vec.p, coeff[2 * (nterms + i)], coeff[2 * (nterms + i) + 1]});
    CU(cudaStreamSynchronize(stream(c)));   // vec.p is reused by the next term
  }
  force::force_chain_w(x, b, level2_coeff, /*naik_in_oprod=*/true, num_naik_terms > 0);
  UploadStream up;
  if (overlap) {
    // The W-level chain (three fifths of the device time) is queued; V and U (two thirds of the bytes that come up)
    // travel on a second stream while it runs, and the V/U-level chain waits for them on the device.
    CU(cudaStreamCreateWithFlags(&up.s, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&up.done, cudaEventDisableTiming));
    CHK(field_to_dev(c, up.s, b.V, b.fs, vlink, host_prec));
    CHK(field_to_dev(c, up.s, b.U, b.fs, ulink, host_prec));
    CU(cudaEventRecord(up.done, up.s));
    CU(cudaStreamWaitEvent(stream(c), up.done, 0));
  }
  force::force_chain_vu(x, b, fat7_coeff, force_filter);
  if (host_prec == 2) x.run(4 * n, force::MomSite<double>{b.U, b.gU, (double *)mom.p, eps, b.fs, n});
  else x.run(4 * n, force::MomSite<float>{b.U, b.gU, (float *)mom.p, eps, b.fs, n});
  CHK(check_launch("fermion force"));
  CHK(d2h(c, momentum, mom.p, (size_t)n * 40 * (host_prec == 2 ? 8 : 4)));
  CU(cudaStreamSynchronize(stream(c)));
  return check_launch("fermion force");
}
