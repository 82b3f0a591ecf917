// tests/host/deflate_host.cu -- TEST INFRASTRUCTURE.  Runs the __host__ __device__ site routines of
// milc_qcd_b200/csrc/deflate.cuh in host loops with the kernels' structure (chunks of sites per "CTA",
// chunk sums added in order per vector, then one update per site), so that tests/test_deflate_host.py
// can compare them with the CPU oracle (oracle/ks_oracle.c kso_deflate) without a GPU.  Built with
//     nvcc -x cu --shared -Xcompiler -fPIC -o tests/host/libdeflate_host.so tests/host/deflate_host.cu
// Nothing here launches a kernel.
#include <algorithm>
#include <vector>

#include "../../milc_qcd_b200/csrc/deflate.cuh"

using namespace b200ks;

// MILC host order su3_vector[V] (6 reals per site, even sites first) -> one parity as 3 planes of double2
static void to_planes(double2 *dst, const double *src, size_t stride, int vh, int pbit) {
  for (int f = 0; f < vh; f++)
    for (int c = 0; c < 3; c++) {
      const double *s = src + 6 * ((size_t)pbit * vh + f) + 2 * c;
      dst[(size_t)c * stride + f] = make_double2(s[0], s[1]);
    }
}

extern "C" void deflate_host(int vol, int nvecs, const double *eigvec, const double *eigval, const double *src, double *dst,
                             double mass, int pbit, int nchunks) {
  const int vh = vol / 2;
  const size_t stride = (size_t)vh + 3;   // deliberately not a multiple of anything
  std::vector<double2> store((size_t)(nvecs + 2) * 3 * stride);
  std::vector<const double2 *> vecs(nvecs);
  for (int j = 0; j < nvecs; j++) {
    double2 *v = store.data() + (size_t)j * 3 * stride;
    to_planes(v, eigvec + (size_t)j * vol * 6, stride, vh, pbit);
    vecs[j] = v;
  }
  double2 *s = store.data() + (size_t)nvecs * 3 * stride, *d = s + 3 * stride;
  to_planes(s, src, stride, vh, pbit);
  to_planes(d, dst, stride, vh, pbit);
  const int per = (vh + nchunks - 1) / nchunks;
  std::vector<double> part((size_t)nvecs * nchunks * 4, 0.0);
  for (int b = 0; b < nchunks; b++) {   // eig_dot_kernel
    const int lo = b * per, hi = std::min(vh, lo + per);
    for (int j = 0; j < nvecs; j++) {
      double acc[4] = {0, 0, 0, 0};
      for (int f = lo; f < hi; f++) eig_dot_site(vecs[j], s, d, stride, f, acc);
      for (int k = 0; k < 4; k++) part[((size_t)j * nchunks + b) * 4 + k] = acc[k];
    }
  }
  std::vector<double2> coef(nvecs);
  for (int j = 0; j < nvecs; j++) {     // eig_coef_kernel
    double t[4] = {0, 0, 0, 0};
    for (int b = 0; b < nchunks; b++)
      for (int k = 0; k < 4; k++) t[k] += part[((size_t)j * nchunks + b) * 4 + k];
    coef[j] = eig_coef(t, eigval[j] + 4.0 * mass * mass);
  }
  for (int f = 0; f < vh; f++) eig_axpy_site(vecs.data(), coef.data(), nvecs, d, stride, f);   // eig_axpy_kernel
  for (int f = 0; f < vh; f++)
    for (int c = 0; c < 3; c++) {
      double *o = dst + 6 * ((size_t)pbit * vh + f) + 2 * c;
      o[0] = d[(size_t)c * stride + f].x;
      o[1] = d[(size_t)c * stride + f].y;
    }
}
