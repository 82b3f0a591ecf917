// deflate.cuh -- low-mode deflation of a trial solution with eigenvectors resident in HBM
// (SURVEY.md section 8 row f4).
//
// The reference (deflate() + project_out(), generic_ks/mat_invert.c:131-183) gives the CG of
// mat_invert_uml_field / mat_invert_cg_field (:186-257,328-402) the exact solution in the span of
// Num low modes of -D_eo D_oe as its starting point: on the sites of one parity
//     dst <- dst - sum_j v_j <v_j|dst> + sum_j v_j <v_j|src> / (lambda_j + 4 m^2) .
// On the host that is 2 Num sweeps over Num vectors (0.5 s for 500 modes at 32^3 x 64, more than the
// whole solve takes here); with the vectors resident (50 MB per vector and parity at 32^3 x 64:
// 500 modes = 50 GB of the 180 GB) it is two passes over them at HBM speed:
//   eig_dot_kernel   every CTA (8 per SM) owns a contiguous chunk of sites and walks the vectors: per vector
//                    the four sums re/im <v|src>, re/im <v|dst> over its chunk -> partials[j][chunk][4]
//   eig_coef_kernel  one thread per vector adds its chunks in order (deterministic) and forms
//                    c_j = <v_j|src>/(lambda_j + 4 m^2) - <v_j|dst>
//   eig_axpy_kernel  dst(x) += sum_j c_j v_j(x), one thread per site
// Algorithmic bytes: 2 x 48 B per site and vector (HBM-bound; src/dst chunks stay in L2).
// The reference removes the modes from dst one after the other (modified Gram-Schmidt order); the
// batch form here is the same for orthonormal vectors and differs by (orthonormality error) x |dst|
// otherwise -- it is a trial solution, the CG that follows corrects either.
//
// The per-site arithmetic is __host__ __device__ (tests/host/deflate_host.cu runs it in host loops
// against the CPU oracle, oracle/ks_oracle.c kso_deflate).
#pragma once
#include "common.cuh"

namespace b200ks {

// acc += { re <v|a>, im <v|a>, re <v|b>, im <v|b> } at site f   (<v|a> = sum_c conj(v_c) a_c)
__host__ __device__ inline void eig_dot_site(const double2 *v, const double2 *a, const double2 *b, size_t stride, int f,
                                             double (&acc)[4]) {
  for (int c = 0; c < 3; c++) {
    const double2 vv = v[(size_t)c * stride + f], aa = a[(size_t)c * stride + f], bb = b[(size_t)c * stride + f];
    acc[0] += vv.x * aa.x + vv.y * aa.y;
    acc[1] += vv.x * aa.y - vv.y * aa.x;
    acc[2] += vv.x * bb.x + vv.y * bb.y;
    acc[3] += vv.x * bb.y - vv.y * bb.x;
  }
}

// c_j from the four sums s = { <v|src>, <v|dst> } and den = lambda_j + 4 m^2
__host__ __device__ inline double2 eig_coef(const double (&s)[4], double den) {
  return make_double2(s[0] / den - s[2], s[1] / den - s[3]);
}

// dst(f) += sum_j coef_j v_j(f)
__host__ __device__ inline void eig_axpy_site(const double2 *const *vecs, const double2 *coef, int nvecs, double2 *dst,
                                              size_t stride, int f) {
  double2 d[3];
  for (int c = 0; c < 3; c++) d[c] = dst[(size_t)c * stride + f];
  for (int j = 0; j < nvecs; j++) {
    const double2 cj = coef[j];
    const double2 *v = vecs[j];
    for (int c = 0; c < 3; c++) {
      const double2 vv = v[(size_t)c * stride + f];
      d[c].x += cj.x * vv.x - cj.y * vv.y;
      d[c].y += cj.x * vv.y + cj.y * vv.x;
    }
  }
  for (int c = 0; c < 3; c++) dst[(size_t)c * stride + f] = d[c];
}

#ifdef __CUDACC__
// CTA b covers the sites [b*per, min(n, (b+1)*per)) for every vector.  kG vectors per pass: the loads of kG
// vectors are in flight together and src/dst are read once per pass (one vector at a time ran at 2.4 TB/s:
// seven sites per thread and a barrier per vector left the memory system idle most of the time).
constexpr int kEigGroup = 4;
__global__ void __launch_bounds__(kBlock)
eig_dot_kernel(const double2 *const *vecs, int nvecs, const double2 *src, const double2 *dst, int stride, int n, int per,
               double *partials) {
  const int lo = blockIdx.x * per, hi = min(n, lo + per);
  for (int j0 = 0; j0 < nvecs; j0 += kEigGroup) {
    double acc[kEigGroup][4];
    const double2 *v[kEigGroup];
#pragma unroll
    for (int g = 0; g < kEigGroup; g++) {
      v[g] = vecs[min(j0 + g, nvecs - 1)];
#pragma unroll
      for (int k = 0; k < 4; k++) acc[g][k] = 0.0;
    }
    for (int f = lo + threadIdx.x; f < hi; f += kBlock) {
#pragma unroll
      for (int g = 0; g < kEigGroup; g++) eig_dot_site(v[g], src, dst, (size_t)stride, f, acc[g]);
    }
#pragma unroll
    for (int g = 0; g < kEigGroup; g++) {
      if (j0 + g < nvecs) block_partials<4>(acc[g], partials + (size_t)(j0 + g) * gridDim.x * 4);
      __syncthreads();   // block_partials' shared scratch is reused by the next vector
    }
  }
}

// one warp per vector adds its chunks (lane l: chunks l, l+32, ...; then the warp tree: a fixed order) and forms
// c_j = <v_j|src>/(lambda_j + 4 m^2) - <v_j|dst>
__global__ void __launch_bounds__(128)
eig_coef_kernel(const double *partials, int nchunks, const double *eigval, double four_m2, int nvecs, double2 *coef) {
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (j >= nvecs) return;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int b = lane; b < nchunks; b += 32)
#pragma unroll
    for (int k = 0; k < 4; k++) s[k] += partials[((size_t)j * nchunks + b) * 4 + k];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if (lane == 0) coef[j] = eig_coef(s, eigval[j] + four_m2);
}

__global__ void __launch_bounds__(kBlock)
eig_axpy_kernel(const double2 *const *vecs, const double2 *coef, int nvecs, double2 *dst, int stride, int n) {
  const int f = blockIdx.x * kBlock + threadIdx.x;
  if (f < n) eig_axpy_site(vecs, coef, nvecs, dst, (size_t)stride, f);
}
#endif

}  // namespace b200ks
