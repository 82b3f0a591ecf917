// internal.h -- what the translation units of libb200ks share on the host side.  The context
// itself is private to b200ks.cu; other units reach the few members they need through these
// accessors.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <string>

#include "../../include/b200ks.h"
#include "common.cuh"

namespace b200ks_host {

int fail(int code, const std::string &msg);          // records the message, returns the code
int check_launch(const char *what);                  // cudaGetLastError -> B200KS_ECUDA
int dev_alloc(b200ks_ctx *c, void **p, size_t bytes);  // cudaMalloc + bookkeeping (b200ks_device_bytes)
void dev_release(b200ks_ctx *c, void *p, size_t bytes);
int stage_get(b200ks_ctx *c, size_t bytes, void **out);   // persistent re-layout staging buffer
// host <-> device copies that bounce pageable host arrays through pinned buffers (b200ks.cu);
// h2d is stream-ordered and returns once the host array has been read, d2h returns when it is complete
int h2d(b200ks_ctx *c, void *dst, const void *src, size_t bytes);
int d2h(b200ks_ctx *c, void *dst, const void *src, size_t bytes);
int h2d_on(b200ks_ctx *c, cudaStream_t on, void *dst, const void *src, size_t bytes);   // h2d on another stream of the device
const b200ks::Geom &geom(const b200ks_ctx *c);
cudaStream_t stream(const b200ks_ctx *c);
int device(const b200ks_ctx *c);
bool partitioned(const b200ks_ctx *c);               // one-rank-per-GPU context with a split direction
// The context an operation without a partitioned implementation (link construction, fermion force) runs on:
// c itself, or -- for the leader of a single-process multi-GPU context -- a full-lattice context on
// its first device, created on first use.  nullptr + error message on failure.
b200ks_ctx *single_gpu_ctx(b200ks_ctx *c);
void count_launch(b200ks_ctx *c);
void *&link_work(b200ks_ctx *c);                     // slot owned by fermion_links.cu
void fermion_links_release(b200ks_ctx *c);           // fermion_links.cu; called by b200ks_destroy

inline int nblocks(int n) { return (n + b200ks::kBlock - 1) / b200ks::kBlock; }

// Launch order of the full-lattice gather kernels (link construction, fermion force): 1 = the two parities interleaved
// CTA by CTA (common.cuh interleaved_site), 0 = all even sites, then all odd ones.  Read per call: B200KS_SITE_ORDER
// is an A/B switch.  Measured at 32^3 x 64 (profiles/force_ab_r02u.json, ncu_staple_r02u.txt): DRAM traffic of a staple
// pass of the link construction 867 -> 709 B per site, 270 -> 236 us, the whole chain 41.4 -> 36.7 ms; backward staple
// pass of the force 1446 -> 1007 B per site, 930 -> 843 us.
constexpr int kDefaultSiteOrder = 1;
inline int site_order() {
  const char *e = getenv("B200KS_SITE_ORDER");
  return e ? (atoi(e) != 0 ? 1 : 0) : kDefaultSiteOrder;
}

}  // namespace b200ks_host

#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return b200ks_host::fail(B200KS_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)
#define CHK(call)          \
  do {                     \
    int r_ = (call);       \
    if (r_ < 0) return r_; \
  } while (0)
