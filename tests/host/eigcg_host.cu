// tests/host/eigcg_host.cu -- TEST INFRASTRUCTURE: the host-side dense algebra of the eigCG solver
// (milc_qcd_b200/csrc/eigcg.cuh namespace dense) behind a C face, so that tests/test_eigcg_host.py can check it
// against numpy without a GPU.
#include <algorithm>
#include <cmath>
#include "../../milc_qcd_b200/csrc/eigcg.cuh"
using b200ks::dense::cd;
extern "C" {
void host_heev(int n, const double *A, double *w, double *Z) {
  std::vector<cd> a((size_t)n * n), z;
  std::vector<double> ww;
  for (size_t k = 0; k < a.size(); k++) a[k] = cd(A[2 * k], A[2 * k + 1]);
  b200ks::dense::heev(n, a, ww, z);
  for (int i = 0; i < n; i++) w[i] = ww[i];
  for (size_t k = 0; k < z.size(); k++) { Z[2 * k] = z[k].real(); Z[2 * k + 1] = z[k].imag(); }
}
void host_orthonormalize(int nrow, int ncol, double *Y) {
  std::vector<cd> y((size_t)nrow * ncol);
  for (size_t k = 0; k < y.size(); k++) y[k] = cd(Y[2 * k], Y[2 * k + 1]);
  b200ks::dense::orthonormalize(nrow, ncol, y);
  for (size_t k = 0; k < y.size(); k++) { Y[2 * k] = y[k].real(); Y[2 * k + 1] = y[k].imag(); }
}
int host_posv(int n, const double *A, double *b) {
  std::vector<cd> a((size_t)n * n), x(n);
  for (size_t k = 0; k < a.size(); k++) a[k] = cd(A[2 * k], A[2 * k + 1]);
  for (int k = 0; k < n; k++) x[k] = cd(b[2 * k], b[2 * k + 1]);
  const bool ok = b200ks::dense::posv(n, a, x);
  for (int k = 0; k < n; k++) { b[2 * k] = x[k].real(); b[2 * k + 1] = x[k].imag(); }
  return ok ? 0 : -1;
}
}
