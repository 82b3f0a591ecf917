// layout.cuh -- MILC host link layout <-> device layout (shared by the solver library and the
// fermion-link construction).
#pragma once
#include "common.cuh"

namespace b200ks {

// Links: host su3_matrix[4*V] as [site][dir][9 complex]  ->  device [dir][9][site].
// Runs once per gauge field; each thread walks its site's 576 contiguous bytes (the
// strided reads are absorbed by L1/L2), the SoA writes are coalesced.
template <typename T, typename TH>
__global__ void __launch_bounds__(kBlock)
pack_link_kernel(typename Vec2<T>::type *d, const TH *h, int lstride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const TH *s = h + (size_t)72 * i;
#pragma unroll 6
  for (int m = 0; m < 36; m++) {  // m = dir*9 + e
    typename Vec2<T>::type o;
    o.x = (T)s[2 * m];
    o.y = (T)s[2 * m + 1];
    d[(size_t)m * lstride + i] = o;
  }
}

// device [dir][9][site] -> host su3_matrix[4*V] layout (inverse of pack_link_kernel)
template <typename T, typename TH>
__global__ void __launch_bounds__(kBlock)
unpack_link_kernel(TH *h, const typename Vec2<T>::type *d, int lstride, int n) {
  const int i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  TH *s = h + (size_t)72 * i;
#pragma unroll 6
  for (int m = 0; m < 36; m++) {
    const auto o = d[(size_t)m * lstride + i];
    s[2 * m] = (TH)o.x;
    s[2 * m + 1] = (TH)o.y;
  }
}

}  // namespace b200ks
