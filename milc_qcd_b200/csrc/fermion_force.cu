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

// one thread per site: the kernel is the __host__ __device__ site functor of force.cuh plus the launch
template <class F>
__global__ void __launch_bounds__(kBlock) force_site_kernel(const F fn, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) fn(i);
}

// the same with a register cap (kMinBlocks 128-thread CTAs per SM: 4 = 128 registers, 3 = 168) for functors that
// declare a static kMinBlocks member
template <class F>
__global__ void __launch_bounds__(kBlock, F::kMinBlocks) force_site_kernel2(const F fn, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) fn(i);
}
template <class F, class = void>
struct WantsCap : std::false_type {};
template <class F>
struct WantsCap<F, std::void_t<decltype(F::kMinBlocks)>> : std::true_type {};

struct DeviceExec {
  b200ks_ctx *c;
  template <class F>
  void run(int n, const F &fn) {
    if constexpr (WantsCap<F>::value) force_site_kernel2<F><<<nblocks(n), kBlock, 0, stream(c)>>>(fn, n);
    else force_site_kernel<F><<<nblocks(n), kBlock, 0, stream(c)>>>(fn, n);
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

// host su3_matrix[4*V] (MILC order) -> full-lattice matrix field (36 planes)
int field_to_dev(b200ks_ctx *c, double2 *dst, size_t fs, const void *host, int host_prec) {
  const int Vh = geom(c).Vh;
  const size_t hs = host_prec == 2 ? 8 : 4;
  const size_t half_bytes = (size_t)Vh * 72 * hs;
  void *stage = nullptr;
  CHK(stage_get(c, half_bytes, &stage));
  for (int p = 0; p < 2; p++) {
    CHK(h2d(c, stage, (const char *)host + (size_t)p * half_bytes, half_bytes));
    if (host_prec == 2) pack_link_kernel<double, double><<<nblocks(Vh), kBlock, 0, stream(c)>>>(dst + (size_t)p * Vh, (const double *)stage, (int)fs, Vh);
    else pack_link_kernel<double, float><<<nblocks(Vh), kBlock, 0, stream(c)>>>(dst + (size_t)p * Vh, (const float *)stage, (int)fs, Vh);
    count_launch(c);
    CU(cudaStreamSynchronize(stream(c)));   // the staging buffer is reused
  }
  return check_launch("pack_link_kernel");
}

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
  {   // A/B switch: the backward staple passes as two kernels with fewer live matrices each (force.cuh)
    const char *e = getenv("B200KS_FORCE_SPLIT");
    b.split = e && atoi(e) != 0;
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
  CHK(field_to_dev(c, b.U, b.fs, ulink, host_prec));
  CHK(field_to_dev(c, b.V, b.fs, vlink, host_prec));
  CHK(field_to_dev(c, b.W, b.fs, wlink, host_prec));
  DeviceExec x{c};
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
  force::force_chain(x, b, fat7_coeff, level2_coeff, /*naik_in_oprod=*/true, force_filter, num_naik_terms > 0);
  if (host_prec == 2) x.run(4 * n, force::MomSite<double>{b.U, b.gU, (double *)mom.p, eps, b.fs, n});
  else x.run(4 * n, force::MomSite<float>{b.U, b.gU, (float *)mom.p, eps, b.fs, n});
  CHK(check_launch("fermion force"));
  CHK(d2h(c, momentum, mom.p, (size_t)n * 40 * (host_prec == 2 ? 8 : 4)));
  CU(cudaStreamSynchronize(stream(c)));
  return check_launch("fermion force");
}
