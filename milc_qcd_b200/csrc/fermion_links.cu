// fermion_links.cu -- host side of the HISQ / asqtad fermion-link construction (links.cuh;
// SURVEY.md section 8 row f1): buffers, the staple recursion of load_fatlinks_cpu
// (generic_ks/fermion_links_fn_load_milc.c:120-275) as a sequence of launches, and the C ABI entry
// points behind qudaLoadKSLink / qudaLoadUnitarizedLink (generic_ks/fermion_links_fn_load_gpu.c).
#include <cstdlib>
#include <cstring>
#include <string>

#include "internal.h"
#include "layout.cuh"
#include "links.cuh"

using namespace b200ks;
using namespace b200ks_host;

#define LAUNCH(c, kern, grid, ...)                                   \
  do {                                                               \
    kern<<<(grid), kBlock, 0, stream(c)>>>(__VA_ARGS__);             \
    count_launch(c);                                                 \
  } while (0)

struct LinkWork {       // full-lattice matrix fields (double), allocated on first use and kept
  double2 *in = nullptr, *v = nullptr, *w = nullptr, *fat = nullptr, *lng = nullptr;   // 36 planes each
  double2 *staple = nullptr, *temp = nullptr;                                           // 9 planes each
  unsigned long long *nsvd = nullptr;
  size_t fstride = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // b200ks_hisq_links_time
};

static ReunitParams reunit_params() {
  // defaults: the reference's RHMC build (ks_imp_rhmc/Make_template:204-206)
  ReunitParams rp;
  rp.allow_svd = getenv("B200KS_REUNIT_ALLOW_SVD") ? atoi(getenv("B200KS_REUNIT_ALLOW_SVD")) : 1;
  rp.svd_rel = getenv("B200KS_REUNIT_SVD_REL_ERROR") ? atof(getenv("B200KS_REUNIT_SVD_REL_ERROR")) : 1e-8;
  rp.svd_abs = getenv("B200KS_REUNIT_SVD_ABS_ERROR") ? atof(getenv("B200KS_REUNIT_SVD_ABS_ERROR")) : 1e-8;
  return rp;
}

static int lw_get(b200ks_ctx *c, LinkWork **out) {
  if (partitioned(c)) return fail(B200KS_ESTATE, "link construction: single-GPU contexts only");
  if (!link_work(c)) {
    LinkWork *w = new LinkWork;
    const size_t V = 2 * (size_t)geom(c).Vh;
    w->fstride = (V + 63) / 64 * 64;
    const size_t m4 = 36 * w->fstride * sizeof(double2), m1 = 9 * w->fstride * sizeof(double2);
    int r = 0;
    r = r < 0 ? r : dev_alloc(c, (void **)&w->in, m4);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->v, m4);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->w, m4);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->fat, m4);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->lng, m4);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->staple, m1);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->temp, m1);
    r = r < 0 ? r : dev_alloc(c, (void **)&w->nsvd, sizeof(unsigned long long));
    if (r < 0) {
      cudaFree(w->in); cudaFree(w->v); cudaFree(w->w); cudaFree(w->fat); cudaFree(w->lng);
      cudaFree(w->staple); cudaFree(w->temp); cudaFree(w->nsvd);
      delete w;
      return r;
    }
    if (cudaEventCreate(&w->ev0) != cudaSuccess || cudaEventCreate(&w->ev1) != cudaSuccess) {
      delete w;
      return fail(B200KS_ECUDA, "cudaEventCreate failed");
    }
    link_work(c) = w;
  }
  *out = (LinkWork *)link_work(c);
  return 0;
}
void b200ks_host::fermion_links_release(b200ks_ctx *c) {
  LinkWork *w = (LinkWork *)link_work(c);
  if (!w) return;
  cudaFree(w->in); cudaFree(w->v); cudaFree(w->w); cudaFree(w->fat); cudaFree(w->lng);
  cudaFree(w->staple); cudaFree(w->temp); cudaFree(w->nsvd);
  if (w->ev0) cudaEventDestroy(w->ev0);
  if (w->ev1) cudaEventDestroy(w->ev1);
  delete w;
  link_work(c) = nullptr;
}

// host su3_matrix[4*V] (MILC order) <-> full-lattice matrix fields
static int links_to_dev(b200ks_ctx *c, const LinkWork &w, double2 *dst, const void *host, int host_prec) {
  const size_t hs = host_prec == 2 ? 8 : 4;
  const size_t half_bytes = (size_t)geom(c).Vh * 72 * hs;
  void *stage = nullptr;
  CHK(stage_get(c, half_bytes, &stage));
  for (int p = 0; p < 2; p++) {
    CHK(h2d(c, stage, (const char *)host + (size_t)p * half_bytes, half_bytes));
    if (host_prec == 2) LAUNCH(c, (pack_link_kernel<double, double>), nblocks(geom(c).Vh), dst + (size_t)p * geom(c).Vh, (const double *)stage, (int)w.fstride, geom(c).Vh);
    else LAUNCH(c, (pack_link_kernel<double, float>), nblocks(geom(c).Vh), dst + (size_t)p * geom(c).Vh, (const float *)stage, (int)w.fstride, geom(c).Vh);
    CU(cudaStreamSynchronize(stream(c)));   // the staging buffer is reused for the next half
  }
  return check_launch("pack_link_kernel");
}
static int links_to_host(b200ks_ctx *c, const LinkWork &w, const double2 *src, void *host, int host_prec) {
  const size_t hs = host_prec == 2 ? 8 : 4;
  const size_t half_bytes = (size_t)geom(c).Vh * 72 * hs;
  void *stage = nullptr;
  CHK(stage_get(c, half_bytes, &stage));
  for (int p = 0; p < 2; p++) {
    if (host_prec == 2) LAUNCH(c, (unpack_link_kernel<double, double>), nblocks(geom(c).Vh), (double *)stage, src + (size_t)p * geom(c).Vh, (int)w.fstride, geom(c).Vh);
    else LAUNCH(c, (unpack_link_kernel<double, float>), nblocks(geom(c).Vh), (float *)stage, src + (size_t)p * geom(c).Vh, (int)w.fstride, geom(c).Vh);
    CHK(d2h(c, (char *)host + (size_t)p * half_bytes, stage, half_bytes));
  }
  return check_launch("unpack_link_kernel");
}

// fat (and lng unless null) from `links`: load_fatlinks_cpu + load_lnglinks on the device.
// coeffs = {one_link, naik, three_staple, five_staple, seven_staple, lepage}
static int smear_dev(b200ks_ctx *c, const LinkWork &w, const double *coeffs, const double2 *links, double2 *fat, double2 *lng) {
  const Geom &g = geom(c);
  const int V = 2 * g.Vh;
  const int grid = nblocks(V);
  const int il = site_order();                                 // staple passes: parities interleaved CTA by CTA
  const int sgrid = il ? interleaved_blocks(g.Vh) : grid;
  const double one_link = coeffs[0], naik = coeffs[1], three = coeffs[2], five = coeffs[3], seven = coeffs[4], lepage = coeffs[5];
  LAUNCH(c, onelink_kernel, grid, fat, links, one_link - 6.0 * lepage, w.fstride, V);
  if (!(three == 0.0 && lepage == 0.0 && five == 0.0)) {
    for (int dir = 0; dir < 4; dir++)
      for (int nu = 0; nu < 4; nu++) {
        if (nu == dir) continue;
        LAUNCH(c, (staple_kernel<true>), sgrid, w.staple, links + (size_t)dir * 9 * w.fstride, links, fat, dir, nu, three, g, w.fstride, V, il);
        if (lepage != 0.0)   // (a zero coefficient adds nothing: the reference computes it anyway)
          LAUNCH(c, (staple_kernel<false>), sgrid, (double2 *)nullptr, w.staple, links, fat, dir, nu, lepage, g, w.fstride, V, il);
        for (int rho = 0; rho < 4; rho++) {
          if (rho == dir || rho == nu) continue;
          LAUNCH(c, (staple_kernel<true>), sgrid, w.temp, w.staple, links, fat, dir, rho, five, g, w.fstride, V, il);
          for (int sig = 0; sig < 4; sig++) {
            if (sig == dir || sig == nu || sig == rho) continue;
            LAUNCH(c, (staple_kernel<false>), sgrid, (double2 *)nullptr, w.temp, links, fat, dir, sig, seven, g, w.fstride, V, il);
          }
        }
      }
  }
  if (lng) LAUNCH(c, longlink_kernel, nblocks(4 * V), lng, links, naik, g, w.fstride, V);
  return check_launch("link smearing");
}

static int unitarize_dev(b200ks_ctx *c, const LinkWork &w, const double2 *V, double2 *W, long long *nsvd) {
  const int n = 2 * geom(c).Vh;
  CU(cudaMemsetAsync(w.nsvd, 0, sizeof(unsigned long long), stream(c)));
  LAUNCH(c, unitarize_kernel, nblocks(4 * n), W, V, w.fstride, n, reunit_params(), w.nsvd);
  unsigned long long h = 0;
  CU(cudaMemcpyAsync(&h, w.nsvd, sizeof(h), cudaMemcpyDeviceToHost, stream(c)));
  CU(cudaStreamSynchronize(stream(c)));
  if (nsvd) *nsvd = (long long)h;
  return check_launch("unitarize_kernel");
}

static int check_link_args(b200ks_ctx *c, const double *coeffs, int host_prec) {
  if (!c || !coeffs) return fail(B200KS_EINVAL, "link construction: null argument");
  if (host_prec != 1 && host_prec != 2) return fail(B200KS_EINVAL, "host_prec must be 1 or 2");
  CU(cudaSetDevice(device(c)));
  return 0;
}

extern "C" int b200ks_ks_links(b200ks_ctx *c, const double *path_coeff, const void *inlink, void *fatlink, void *longlink,
                               int host_prec) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  CHK(check_link_args(c, path_coeff, host_prec));
  if (!inlink || !fatlink) return fail(B200KS_EINVAL, "b200ks_ks_links: null field");
  LinkWork *w = nullptr;
  CHK(lw_get(c, &w));
  CHK(links_to_dev(c, *w, w->in, inlink, host_prec));
  CHK(smear_dev(c, *w, path_coeff, w->in, w->fat, longlink ? w->lng : nullptr));
  CHK(links_to_host(c, *w, w->fat, fatlink, host_prec));
  if (longlink) CHK(links_to_host(c, *w, w->lng, longlink, host_prec));
  return 0;
}

extern "C" int b200ks_unitarized_links(b200ks_ctx *c, const double *path_coeff, const void *inlink, void *vlink, void *wlink,
                                       int host_prec, long long *nsvd) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  CHK(check_link_args(c, path_coeff, host_prec));
  if (!inlink || !wlink) return fail(B200KS_EINVAL, "b200ks_unitarized_links: null field");
  LinkWork *w = nullptr;
  CHK(lw_get(c, &w));
  CHK(links_to_dev(c, *w, w->in, inlink, host_prec));
  CHK(smear_dev(c, *w, path_coeff, w->in, w->v, nullptr));
  CHK(unitarize_dev(c, *w, w->v, w->w, nsvd));
  if (vlink) CHK(links_to_host(c, *w, w->v, vlink, host_prec));
  CHK(links_to_host(c, *w, w->w, wlink, host_prec));
  return 0;
}

// the whole chain with the intermediate fields resident: one upload, two (to four) downloads
extern "C" int b200ks_hisq_links(b200ks_ctx *c, const double *coeff1, const double *coeff2, const void *inlink, void *vlink,
                                 void *wlink, void *fatlink, void *longlink, int host_prec, long long *nsvd) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  CHK(check_link_args(c, coeff1, host_prec));
  if (!coeff2 || !inlink) return fail(B200KS_EINVAL, "b200ks_hisq_links: null argument");
  LinkWork *w = nullptr;
  CHK(lw_get(c, &w));
  CHK(links_to_dev(c, *w, w->in, inlink, host_prec));
  CHK(smear_dev(c, *w, coeff1, w->in, w->v, nullptr));
  CHK(unitarize_dev(c, *w, w->v, w->w, nsvd));
  CHK(smear_dev(c, *w, coeff2, w->w, w->fat, w->lng));
  if (vlink) CHK(links_to_host(c, *w, w->v, vlink, host_prec));
  if (wlink) CHK(links_to_host(c, *w, w->w, wlink, host_prec));
  if (fatlink) CHK(links_to_host(c, *w, w->fat, fatlink, host_prec));
  if (longlink) CHK(links_to_host(c, *w, w->lng, longlink, host_prec));
  return 0;
}

// Benchmark face: Haar-random thin links generated on the device (seed), the chain run `reps` times
// with everything resident; *ms = CUDA-event milliseconds per chain.  The input and the four outputs
// can be read back with b200ks_hisq_links_fetch for the CPU comparison.
extern "C" int b200ks_hisq_links_time(b200ks_ctx *c, const double *coeff1, const double *coeff2, unsigned long long seed,
                                      int reps, double *ms, long long *nsvd) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  CHK(check_link_args(c, coeff1, 2));
  if (!coeff2 || !ms || reps <= 0) return fail(B200KS_EINVAL, "b200ks_hisq_links_time: bad argument");
  LinkWork *w = nullptr;
  CHK(lw_get(c, &w));
  const int V = 2 * geom(c).Vh;
  LAUNCH(c, synth_thin_kernel, nblocks(V), w->in, geom(c), w->fstride, V, (uint64_t)seed);
  auto chain = [&]() -> int {
    CHK(smear_dev(c, *w, coeff1, w->in, w->v, nullptr));
    CHK(unitarize_dev(c, *w, w->v, w->w, nsvd));
    CHK(smear_dev(c, *w, coeff2, w->w, w->fat, w->lng));
    return 0;
  };
  CHK(chain());
  CU(cudaEventRecord(w->ev0, stream(c)));
  for (int k = 0; k < reps; k++) CHK(chain());
  CU(cudaEventRecord(w->ev1, stream(c)));
  CU(cudaEventSynchronize(w->ev1));
  float t = 0;
  CU(cudaEventElapsedTime(&t, w->ev0, w->ev1));
  *ms = (double)t / reps;
  return check_launch("hisq link chain");
}

// which: 0 input thin links, 1 V, 2 W, 3 fat, 4 long (of the last chain run on this context)
extern "C" int b200ks_hisq_links_fetch(b200ks_ctx *c, int which, void *host, int host_prec) {
  if (c && !(c = single_gpu_ctx(c))) return B200KS_ECUDA;   // (multi-GPU leader: its full-lattice context)
  if (!c || !host || which < 0 || which > 4) return fail(B200KS_EINVAL, "b200ks_hisq_links_fetch: bad argument");
  if (host_prec != 1 && host_prec != 2) return fail(B200KS_EINVAL, "host_prec must be 1 or 2");
  if (!link_work(c)) return fail(B200KS_ESTATE, "no link construction has run on this context");
  CU(cudaSetDevice(device(c)));
  LinkWork *w = (LinkWork *)link_work(c);
  const double2 *src = which == 0 ? w->in : which == 1 ? w->v : which == 2 ? w->w : which == 3 ? w->fat : w->lng;
  return links_to_host(c, *w, src, host, host_prec);
}
