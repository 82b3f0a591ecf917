// tests/host/force_host.cu -- TEST INFRASTRUCTURE.  Runs the __host__ __device__ site routines of
// milc_qcd_b200/csrc/force.cuh (the bodies of the CUDA force kernels and the chain that sequences
// them) in plain host loops, so that tests/test_force_host.py can compare them with the CPU oracle
// (oracle/ks_force_oracle.c) without a GPU.  Built by tests/test_force_host.py with
//     nvcc -x cu --shared -Xcompiler -fPIC -o tests/host/libforce_host.so tests/host/force_host.cu
// Nothing here launches a kernel.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../milc_qcd_b200/csrc/force.cuh"

using namespace b200ks::force;

struct HostExec {
  template <class F>
  void run(int n, const F &f) {
    for (int i = 0; i < n; i++) f(i);
  }
};

// MILC host links su3_matrix[4*V] ([site][dir][9][2]) -> 36 planes of double2
static void to_planes(double2 *dst, const double *src, size_t fs, int nsites) {
  for (int f = 0; f < nsites; f++)
    for (int m = 0; m < 36; m++) dst[(size_t)m * fs + f] = make_double2(src[(size_t)72 * f + 2 * m], src[(size_t)72 * f + 2 * m + 1]);
}

extern "C" void force_host(const int *dims, const double *coeffs1, const double *coeffs2, const double *U, const double *V,
                           const double *W, const double *multi_x, const double *c1, const double *c3, int nterms,
                           int n_naik_terms, double eps, int naik_in_oprod, double filter, int split, double *mom) {
  ForceBufs b;
  b.split = split == 1;   // 1: the four-kernel form of the backward staple passes, 2: the two-role form (StapleBwdPairSite)
  b.pair = split == 2 ? 4 : split == 3 ? 3 : 0;
  for (int d = 0; d < 4; d++) b.g.L[d] = dims[d];
  const int n = dims[0] * dims[1] * dims[2] * dims[3];
  b.g.Vh = n / 2;
  b.nsites = n;
  b.fs = (size_t)n + 5;   // deliberately not a multiple of anything
  std::vector<double2> store((size_t)(7 * 36 + 4 * 9) * b.fs, make_double2(0.0, 0.0));
  double2 *p = store.data();
  b.U = p; p += 36 * b.fs;
  b.V = p; p += 36 * b.fs;
  b.W = p; p += 36 * b.fs;
  b.gfat = p; p += 36 * b.fs;
  b.glng = p; p += 36 * b.fs;
  b.gW = p; p += 36 * b.fs;
  b.gU = p; p += 36 * b.fs;
  b.st3 = p; p += 9 * b.fs;
  b.st5 = p; p += 9 * b.fs;
  b.g3 = p; p += 9 * b.fs;
  b.g5 = p; p += 9 * b.fs;
  to_planes(b.U, U, b.fs, n);
  to_planes(b.V, V, b.fs, n);
  to_planes(b.W, W, b.fs, n);
  HostExec x;
  for (int j = 0; j < nterms; j++) x.run(n, OprodSite{b.g, b.gfat, b.glng, b.fs, multi_x + (size_t)j * n * 6, c1[j], c3[j]});
  // the terms solved with a Naik epsilon: the last n_naik_terms fields, weights c1[nterms + i], c3[nterms + i]
  for (int i = 0; i < n_naik_terms; i++)
    x.run(n, OprodSite{b.g, b.gW, b.gU, b.fs, multi_x + (size_t)(nterms - n_naik_terms + i) * n * 6, c1[nterms + i], c3[nterms + i]});
  force_chain(x, b, coeffs1, coeffs2, naik_in_oprod != 0, filter, n_naik_terms > 0);
  x.run(4 * n, MomSite<double>{b.U, b.gU, mom, eps, b.fs, n});
}

// launch order of the full-lattice gather kernels (common.cuh): out[i] = site of launch index i (-1 past the end),
// i < interleaved_blocks(Vh) * kBlock; returns that count
#include "../../milc_qcd_b200/csrc/common.cuh"
extern "C" int interleaved_order(int Vh, int *out, int cap) {
  const int n = b200ks::interleaved_blocks(Vh) * b200ks::kBlock;
  for (int i = 0; i < n && i < cap; i++) out[i] = b200ks::interleaved_site(i, Vh);
  return n;
}
