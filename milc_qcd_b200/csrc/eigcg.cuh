// eigcg.cuh -- kernels and small dense algebra of the eigCG / incremental eigCG solver
// (generic_ks/inc_eigcg.c, A. Stathopoulos and K. Orginos, arXiv:0707.0131).
//
// eigCG is the plain CG for (4 m^2 - D^2) x = b plus bookkeeping: the normalised residuals of the last m
// iterations are kept as a search space V (they are the Lanczos vectors of the same Krylov space), the
// Lanczos matrix T = V^+ A V comes for free from the CG coefficients, and whenever the window is full it is
// compressed to the 2 Nvecs Ritz vectors of T_m and T_{m-1}.  The reference does the window on the host
// (m + Nvecs_max vectors of 50 MB at 32^3 x 64 swept by one core); here the window lives in HBM next to the
// deflation set, one parity half per vector, and three kinds of kernels touch it:
//   eig_dot_kernel / eig_axpy_kernel (deflate.cuh)   n dot products against one vector / one combination
//   eig_rotate_kernel                                V' = V C for a small coefficient matrix C (window restart,
//                                                    Rayleigh-Ritz): every input is read once per 8 outputs
//   eig_dotsum_kernel                                chunk partial sums -> n complex numbers
// The dense problems (m x m Hermitian eigenproblems, a QR of m x 2 Nvecs, a Cholesky solve) are a few hundred
// kflop each and stay on the host: Jacobi rotations, modified Gram-Schmidt, Cholesky -- no LAPACK dependency.
#pragma once
#include <complex>
#include <vector>

#include "common.cuh"

namespace b200ks {

// out[j] = { <v_j|a>, <v_j|b> } from eig_dot_kernel's partials [j][chunk][4]; one warp per vector, fixed order
__global__ void __launch_bounds__(128) eig_dotsum_kernel(const double *partials, int nchunks, int nvecs, double *out) {
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
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 4; k++) out[4 * j + k] = s[k];
}

// out_o(f) = sum_jj coef[jj * nout + o] * in_jj(f) for the 8 outputs o = 8*blockIdx.y .. ; one thread per site
constexpr int kRotOut = 8;
__global__ void __launch_bounds__(kBlock)
eig_rotate_kernel(const double2 *const *in, int nin, double2 *const *out, int nout, const double2 *coef, int stride, int n) {
  const int f = blockIdx.x * kBlock + threadIdx.x;
  if (f >= n) return;
  const int o0 = blockIdx.y * kRotOut;
  double2 acc[kRotOut][3];
#pragma unroll
  for (int o = 0; o < kRotOut; o++)
#pragma unroll
    for (int c = 0; c < 3; c++) acc[o][c] = make_double2(0.0, 0.0);
  for (int jj = 0; jj < nin; jj++) {
    const double2 *v = in[jj];
    double2 x[3];
#pragma unroll
    for (int c = 0; c < 3; c++) x[c] = v[(size_t)c * stride + f];
#pragma unroll
    for (int o = 0; o < kRotOut; o++) {
      const double2 cc = (o0 + o < nout) ? __ldg(&coef[(size_t)jj * nout + o0 + o]) : make_double2(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        acc[o][c].x += cc.x * x[c].x - cc.y * x[c].y;
        acc[o][c].y += cc.x * x[c].y + cc.y * x[c].x;
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kRotOut; o++)
    if (o0 + o < nout)
#pragma unroll
      for (int c = 0; c < 3; c++) out[o0 + o][(size_t)c * stride + f] = acc[o][c];
}

// ---- host: small dense complex algebra -------------------------------------------------------------------
namespace dense {
using cd = std::complex<double>;

// Hermitian eigenproblem A = Z diag(w) Z^+, A n x n row-major (only its Hermitian part is used), w ascending,
// eigenvectors in the COLUMNS of Z (row-major n x n).  Cyclic Jacobi.
inline void heev(int n, std::vector<cd> A, std::vector<double> &w, std::vector<cd> &Z) {
  Z.assign((size_t)n * n, cd(0, 0));
  for (int i = 0; i < n; i++) {
    Z[(size_t)i * n + i] = 1.0;
    A[(size_t)i * n + i] = A[(size_t)i * n + i].real();
    for (int j = i + 1; j < n; j++) {   // symmetrise from the upper triangle (what LAPACK's uplo = 'U' reads)
      A[(size_t)j * n + i] = std::conj(A[(size_t)i * n + j]);
    }
  }
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++) {
      diag += std::norm(A[(size_t)i * n + i]);
      for (int j = i + 1; j < n; j++) off += std::norm(A[(size_t)i * n + j]);
    }
    if (off <= 1e-32 * (diag + off) || off == 0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const cd apq = A[(size_t)p * n + q];
        const double g = std::abs(apq);
        if (g == 0.0) continue;
        const double app = A[(size_t)p * n + p].real(), aqq = A[(size_t)q * n + q].real();
        const cd ph = apq / g;                       // a_pq = g e^{i phi}
        const double tau = (aqq - app) / (2.0 * g);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = t * cs;
        // columns p, q <- [p q] R with R = [[cs, sn ph], [-sn conj(ph), cs]]  (unitary)
        for (int k = 0; k < n; k++) {
          const cd akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = cs * akp - sn * std::conj(ph) * akq;
          A[(size_t)k * n + q] = sn * ph * akp + cs * akq;
          const cd zkp = Z[(size_t)k * n + p], zkq = Z[(size_t)k * n + q];
          Z[(size_t)k * n + p] = cs * zkp - sn * std::conj(ph) * zkq;
          Z[(size_t)k * n + q] = sn * ph * zkp + cs * zkq;
        }
        for (int k = 0; k < n; k++) {                // rows p, q <- R^+ [p; q]
          const cd apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = cs * apk - sn * ph * aqk;
          A[(size_t)q * n + k] = sn * std::conj(ph) * apk + cs * aqk;
        }
        A[(size_t)p * n + q] = A[(size_t)q * n + p] = 0.0;
        A[(size_t)p * n + p] = A[(size_t)p * n + p].real();
        A[(size_t)q * n + q] = A[(size_t)q * n + q].real();
      }
  }
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int a, int b) { return A[(size_t)a * n + a].real() < A[(size_t)b * n + b].real(); });
  w.resize(n);
  std::vector<cd> Zs((size_t)n * n);
  for (int j = 0; j < n; j++) {
    w[j] = A[(size_t)idx[j] * n + idx[j]].real();
    for (int k = 0; k < n; k++) Zs[(size_t)k * n + j] = Z[(size_t)k * n + idx[j]];
  }
  Z.swap(Zs);
}

// orthonormalise the ncol columns of Y (row-major nrow x ncol), modified Gram-Schmidt twice; a column that
// vanishes is replaced by zero (its Ritz vector then drops out)
inline void orthonormalize(int nrow, int ncol, std::vector<cd> &Y) {
  for (int j = 0; j < ncol; j++) {
    for (int pass = 0; pass < 2; pass++)
      for (int k = 0; k < j; k++) {
        cd d = 0;
        for (int i = 0; i < nrow; i++) d += std::conj(Y[(size_t)i * ncol + k]) * Y[(size_t)i * ncol + j];
        for (int i = 0; i < nrow; i++) Y[(size_t)i * ncol + j] -= d * Y[(size_t)i * ncol + k];
      }
    double nn = 0;
    for (int i = 0; i < nrow; i++) nn += std::norm(Y[(size_t)i * ncol + j]);
    nn = std::sqrt(nn);
    for (int i = 0; i < nrow; i++) Y[(size_t)i * ncol + j] = nn > 1e-14 ? Y[(size_t)i * ncol + j] / nn : cd(0, 0);
  }
}

// solves A x = b for a Hermitian positive definite A (row-major n x n, upper triangle used); false if not p.d.
inline bool posv(int n, std::vector<cd> A, std::vector<cd> &b) {
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) A[(size_t)j * n + i] = std::conj(A[(size_t)i * n + j]);
  // A = L L^+
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j].real();
    for (int k = 0; k < j; k++) d -= std::norm(A[(size_t)j * n + k]);
    if (!(d > 0)) return false;
    d = std::sqrt(d);
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      cd s = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) s -= A[(size_t)i * n + k] * std::conj(A[(size_t)j * n + k]);
      A[(size_t)i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; i++) {          // L y = b
    cd s = b[i];
    for (int k = 0; k < i; k++) s -= A[(size_t)i * n + k] * b[k];
    b[i] = s / A[(size_t)i * n + i].real();
  }
  for (int i = n - 1; i >= 0; i--) {     // L^+ x = y
    cd s = b[i];
    for (int k = i + 1; k < n; k++) s -= std::conj(A[(size_t)k * n + i]) * b[k];
    b[i] = s / A[(size_t)i * n + i].real();
  }
  return true;
}
}  // namespace dense

}  // namespace b200ks
