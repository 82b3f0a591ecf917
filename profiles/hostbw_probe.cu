// hostbw_probe.cu -- what the host side of the drop-in seam has to work with on a given box:
// PCIe copy rates from pinned and from pageable memory, threaded memcpy into a pinned bounce buffer,
// and the content-fingerprint rate.  Build: nvcc -O2 -o profiles/bin/hostbw_probe profiles/hostbw_probe.cu
//   -Iinclude -Lmilc_qcd_b200 -lb200ks -Xlinker -rpath -Xlinker '$ORIGIN/../../milc_qcd_b200'
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "b200ks.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t n = (size_t)512 << 20;
  char *pageable = (char *)malloc(n), *pinned = nullptr, *dev = nullptr;
  memset(pageable, 1, n);
  cudaMallocHost(&pinned, n);
  memset(pinned, 2, n);
  cudaMalloc(&dev, n);
  printf("{\"host_threads\": %u", std::thread::hardware_concurrency());
  for (int rep = 0; rep < 2; rep++) {
    double t = now(); cudaMemcpy(dev, pinned, n, cudaMemcpyHostToDevice); t = now() - t;
    if (rep) printf(", \"h2d_pinned_gbs\": %.2f", n / t / 1e9);
    t = now(); cudaMemcpy(pinned, dev, n, cudaMemcpyDeviceToHost); t = now() - t;
    if (rep) printf(", \"d2h_pinned_gbs\": %.2f", n / t / 1e9);
    t = now(); cudaMemcpy(dev, pageable, n, cudaMemcpyHostToDevice); t = now() - t;
    if (rep) printf(", \"h2d_pageable_gbs\": %.2f", n / t / 1e9);
    t = now(); cudaMemcpy(pageable, dev, n, cudaMemcpyDeviceToHost); t = now() - t;
    if (rep) printf(", \"d2h_pageable_gbs\": %.2f", n / t / 1e9);
  }
  printf(", \"memcpy_pageable_to_pinned_gbs\": {");
  bool first = true;
  for (int nt : {1, 2, 4, 8, 16, 32}) {
    double best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      double t = now();
      std::vector<std::thread> th;
      const size_t per = n / nt;
      for (int k = 0; k < nt; k++) th.emplace_back([=]() { memcpy(pinned + k * per, pageable + k * per, per); });
      for (auto &x : th) x.join();
      t = now() - t;
      if (t < best) best = t;
    }
    printf("%s\"%d\": %.2f", first ? "" : ", ", nt, n / best / 1e9);
    first = false;
  }
  printf("}");
  {
    double best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      double t = now();
      volatile unsigned long long h = b200ks_fingerprint(pageable, n);
      (void)h;
      t = now() - t;
      if (t < best) best = t;
    }
    printf(", \"fingerprint_gbs\": %.2f", n / best / 1e9);
  }
  // cudaHostRegister cost (page-locking MILC's arrays in place)
  {
    double t = now();
    cudaError_t e = cudaHostRegister(pageable, n, cudaHostRegisterDefault);
    t = now() - t;
    printf(", \"host_register_512MB_ms\": %.2f, \"host_register_ok\": %d", t * 1e3, e == cudaSuccess);
    if (e == cudaSuccess) {
      double t2 = now(); cudaMemcpy(dev, pageable, n, cudaMemcpyHostToDevice); t2 = now() - t2;
      printf(", \"h2d_registered_gbs\": %.2f", n / t2 / 1e9);
      t2 = now(); cudaHostUnregister(pageable); t2 = now() - t2;
      printf(", \"host_unregister_ms\": %.2f", t2 * 1e3);
    }
  }
  printf("}\n");
  return 0;
}
