// quda_shim.cu -- route-2 boundary: the quda* entry points MILC's own GPU glue calls
// (include/quda_milc_interface.h), implemented on top of the b200ks C ABI.
//
// Error convention at this boundary is MILC's: print and terminate(1)
// (generic/com_vanilla.c:203-211; e.g. generic_ks/d_congrad5_fn_gpu.c:98-101).  The glue
// gives no way to return an error, so a failed library call ends the run loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200ks.h"
#include "../../include/quda_milc_interface.h"
#include "common.cuh"

namespace {

struct ShimState {
  b200ks_ctx *ctx = nullptr;
  int latsize[4] = {0, 0, 0, 0};
  int device = 0;
  int verbosity = QUDA_SUMMARIZE;
  double force_filter = 0.0;   // qudaHisqParamsInit
  bool inited = false;
} S;

[[noreturn]] void die(const char *where) {
  printf("%s: libb200ks error: %s\n", where, b200ks_last_error());
  printf("Termination: node 0, status = 1\n");
  fflush(stdout);
  exit(1);
}

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s ? atoi(s) : dflt;
}

void ensure_ctx(const char *where) {
  if (!S.inited) {
    printf("%s: qudaInit was not called\n", where);
    exit(1);
  }
  if (!S.ctx) {
    S.ctx = b200ks_create(S.latsize, S.device);
    if (!S.ctx) die(where);
    if (b200ks_num_gpus(S.ctx) > 1 && S.verbosity >= QUDA_SUMMARIZE)
      printf("libb200ks: lattice spread over %d GPUs behind the seam (B200KS_NGPU), first device %d\n", b200ks_num_gpus(S.ctx), S.device);
  }
}

// The QUDA seam signals new links with *num_iters == -1 (d_congrad5_fn_gpu.c:121-126); that
// flag is not set when boundary_twist_fn edits the links in place
// (fermion_links_fn_twist_milc.c:318-400).  B200KS_ALWAYS_RELOAD_LINKS selects what to do about it
// (b200ks_links_sync modes):
//   unset / 2  fingerprint the two host arrays on host threads WHILE the solve runs on the resident
//              links; re-upload and repeat the solve only if they did change;
//   3          fingerprint before the solve (blocking, ~70 ms per call at 32^3x64);
//   1          re-upload on every call;
//   0          trust the flag, as the reference's own QUDA path does.
void ensure_links(const char *where, const void *fat, const void *lng, int ext_prec, int *num_iters) {
  static const int mode = env_int("B200KS_ALWAYS_RELOAD_LINKS", 2);
  const int hint = (num_iters && *num_iters == -1) ? 1 : 0;
  if (b200ks_links_sync(S.ctx, fat, lng, ext_prec, hint, mode < 0 || mode > 3 ? 2 : mode) < 0) die(where);
}

int milc_parity(QudaParity p, const char *where) {
  if (p == QUDA_EVEN_PARITY) return B200KS_EVEN;
  if (p == QUDA_ODD_PARITY) return B200KS_ODD;
  printf("%s: Unrecognised parity\n", where);
  exit(2);
}

// The seam only carries qic->max * qic->nrestart (d_congrad5_fn_gpu.c:104); the CPU
// algorithm needs the two factors.  MILC inputs conventionally use 5 restarts
// (max_cg_restarts 5 in the shipped samples); override with B200KS_NRESTART.
// The product never exceeds the cap the seam was given.
void split_iters(int total, b200ks_invert_args *a) {
  int nr = env_int("B200KS_NRESTART", 5);
  if (nr < 1) nr = 1;
  if (total < nr) nr = 1;
  a->nrestart = nr;
  a->max_iter = total / nr > 0 ? total / nr : 1;
}

__global__ void mom_action_kernel(const char *site, size_t mom_offset, size_t size, long nsites, int prec,
                                  b200ks::ReduceWs ws, double *out) {
  const long i = (long)blockIdx.x * b200ks::kBlock + threadIdx.x;
  double s[1] = {0};
  if (i < nsites) {
    const char *m = site + (size_t)i * size + mom_offset;
    for (int dir = 0; dir < 4; dir++) {
      // anti_hermitmat = {m01, m02, m12 (complex), m00im, m11im, m22im, space} = 10 reals
      double v[9];
      for (int k = 0; k < 9; k++)
        v[k] = (prec == 2) ? ((const double *)m)[dir * 10 + k] : (double)((const float *)m)[dir * 10 + k];
      double sum = 0;
      for (int k = 0; k < 6; k++) sum += v[k] * v[k];
      for (int k = 6; k < 9; k++) sum += 0.5 * v[k] * v[k];
      s[0] += sum - 4.0;
    }
  }
  b200ks::grid_reduce<1>(s, ws, out);
}

}  // namespace

extern "C" {

void qudaInit(QudaInitArgs_t input) {
  if (S.inited) return;
  for (int d = 0; d < 4; d++) S.latsize[d] = input.layout.latsize[d];
  S.device = input.layout.device;
  S.verbosity = (int)input.verbosity;
  // One MILC rank drives all GPUs: the machine grid MILC reports
  // (generic/milc_to_quda_utilities.c:13-36) is that of its MPI ranks and must be 1x1x1x1; the number
  // of devices behind the seam is B200KS_NGPU (b200ks_create -> b200ks_create_multi).
  if (input.layout.machsize)
    for (int d = 0; d < 4; d++)
      if (input.layout.machsize[d] != 1) {
        printf("qudaInit: libb200ks is driven by ONE MILC rank (machine grid must be 1x1x1x1); set B200KS_NGPU=N to "
               "spread the lattice over N GPUs behind it\n");
        exit(1);
      }
  if (b200ks_device_count() < 1) {
    printf("qudaInit: no sm_100 GPU visible; libb200ks has no CPU fallback\n");
    exit(1);
  }
  cudaSetDevice(S.device);
  S.inited = true;
}

void qudaSetMPICommHandle(void *) {}

void qudaFinalize(void) {
  if (S.ctx) b200ks_destroy(S.ctx);
  S = ShimState();
}

void *qudaAllocatePinned(size_t bytes) {
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    printf("qudaAllocatePinned: cudaMallocHost(%zu) failed\n", bytes);
    exit(1);
  }
  return p;
}
void qudaFreePinned(void *p) { cudaFreeHost(p); }
void *qudaAllocateManaged(size_t bytes) { return qudaAllocatePinned(bytes); }
void qudaFreeManaged(void *p) { qudaFreePinned(p); }

void qudaInvert(int external_precision, int quda_precision, double mass, QudaInvertArgs_t inv_args,
                double target_residual, double target_fermilab_residual, const void *const fat,
                const void *const lng, void *source, void *solution, double *const final_residual,
                double *const final_fermilab_residual, int *num_iters) {
  static const char where[] = "qudaInvert";
  (void)quda_precision;
  ensure_ctx(where);
  ensure_links(where, fat, lng, external_precision, num_iters);
  b200ks_invert_args a;
  memset(&a, 0, sizeof(a));
  a.parity = milc_parity(inv_args.evenodd, where);
  split_iters(inv_args.max_iter, &a);
  a.resid = target_residual;
  a.relresid = target_fermilab_residual;
  a.mixed_precision = inv_args.mixed_precision;
  b200ks_invert_result r;
  const int it = b200ks_congrad(S.ctx, source, solution, mass, &a, &r, external_precision);
  if (it < 0) die(where);
  *final_residual = sqrt(r.final_rsq);
  // MILC's glue squares what it gets back (d_congrad5_fn_gpu.c:150-151); the CPU path leaves
  // relative_residue() itself in qic->final_relrsq (d_congrad5_fn_milc.c:37-56,218-221)
  *final_fermilab_residual = sqrt(r.final_relrsq);
  *num_iters = it;
  if (S.verbosity >= QUDA_VERBOSE)
    printf("qudaInvert: %d iterations, true |r|/|b| = %e, %.3e s on device\n", it, *final_residual, r.device_seconds);
}

void qudaInvertMsrc(int external_precision, int quda_precision, double mass, QudaInvertArgs_t inv_args,
                    double target_residual, double target_fermilab_residual, const void *const fat,
                    const void *const lng, void **sourceArray, void **solutionArray, double *const final_residual,
                    double *const final_fermilab_residual, int *num_iters, int num_src) {
  // all sources at once through the multi-right-hand-side CG; reports the worst residual and the
  // total iteration count (what the reference's loop, d_congrad5_fn_milc.c:409-417, returns)
  static const char where[] = "qudaInvertMsrc";
  (void)quda_precision;
  ensure_ctx(where);
  ensure_links(where, fat, lng, external_precision, num_iters);
  b200ks_invert_args a;
  memset(&a, 0, sizeof(a));
  a.parity = milc_parity(inv_args.evenodd, where);
  split_iters(inv_args.max_iter, &a);
  a.resid = target_residual;
  a.relresid = target_fermilab_residual;
  a.mixed_precision = inv_args.mixed_precision;
  std::vector<b200ks_invert_result> r(num_src > 0 ? num_src : 1);
  const int it = b200ks_congrad_block(S.ctx, num_src, (const void *const *)sourceArray, (void *const *)solutionArray, mass, &a,
                                      r.data(), external_precision);
  if (it < 0) die(where);
  double worst = 0, worst_rel = 0;
  for (int k = 0; k < num_src; k++) {
    worst = std::max(worst, sqrt(r[k].final_rsq));
    worst_rel = std::max(worst_rel, sqrt(r[k].final_relrsq));
  }
  *final_residual = worst;
  *final_fermilab_residual = worst_rel;
  *num_iters = it;
  if (S.verbosity >= QUDA_VERBOSE)
    printf("qudaInvertMsrc: %d sources, %d iterations in total, worst true |r|/|b| = %e\n", num_src, it, worst);
}

void qudaMultishiftInvert(int external_precision, int precision, int num_offsets, double *const offset,
                          QudaInvertArgs_t inv_args, const double *target_residual,
                          const double *target_fermilab_residual, const void *const fat, const void *const lng,
                          void *source, void **solutionArray, double *const final_residual,
                          double *const final_fermilab_residual, int *num_iters) {
  static const char where[] = "qudaMultishiftInvert";
  (void)precision;
  (void)target_fermilab_residual;
  ensure_ctx(where);
  ensure_links(where, fat, lng, external_precision, num_iters);
  b200ks_invert_args a;
  memset(&a, 0, sizeof(a));
  a.parity = milc_parity(inv_args.evenodd, where);
  // no restarts in this algorithm: the product is the cap (ks_multicg_offset.c:97)
  a.max_iter = inv_args.max_iter;
  a.nrestart = 1;
  a.resid = target_residual[0];  // convergence is judged on the smallest shift only (:280,381)
  a.relresid = 0;
  a.mixed_precision = inv_args.mixed_precision;
  std::vector<b200ks_invert_result> r(num_offsets > 0 ? num_offsets : 1);
  const int it = b200ks_multicg(S.ctx, source, solutionArray, offset, num_offsets, &a, r.data(), external_precision);
  if (it < 0) die(where);
  for (int j = 0; j < num_offsets; j++) {
    final_residual[j] = sqrt(r[j].final_rsq);
    final_fermilab_residual[j] = 0;
  }
  *num_iters = it;
}

void qudaDslash(int external_precision, int quda_precision, QudaInvertArgs_t inv_args, const void *const fat,
                const void *const lng, void *source, void *solution, int *num_iters) {
  static const char where[] = "qudaDslash";
  (void)quda_precision;
  ensure_ctx(where);
  ensure_links(where, fat, lng, external_precision, num_iters);
  if (b200ks_dslash(S.ctx, source, solution, milc_parity(inv_args.evenodd, where), external_precision) < 0) die(where);
  if (num_iters) *num_iters = 0;
}

void qudaLoadKSLink(int precision, QudaFatLinkArgs_t, const double path_coeff[6], void *inlink, void *fatlink,
                    void *longlink) {
  static const char where[] = "qudaLoadKSLink";
  ensure_ctx(where);
  if (b200ks_ks_links(S.ctx, path_coeff, inlink, fatlink, longlink, precision) < 0) die(where);
}

void qudaLoadUnitarizedLink(int precision, QudaFatLinkArgs_t, const double path_coeff[6], void *inlink, void *fatlink,
                            void *ulink) {
  static const char where[] = "qudaLoadUnitarizedLink";
  ensure_ctx(where);
  long long nsvd = 0;
  if (b200ks_unitarized_links(S.ctx, path_coeff, inlink, fatlink, ulink, precision, &nsvd) < 0) die(where);
  if (S.verbosity >= QUDA_VERBOSE) printf("qudaLoadUnitarizedLink: %lld links took the SVD branch\n", nsvd);
}

void qudaHisqParamsInit(QudaHisqParams_t hisq_params) {
  // the force takes W and V from the caller, so the reunitarisation switches have nothing to act
  // on here; the filter goes to every later qudaHisqForce (fermion_force_hisq_multi.c:2258-2266)
  S.force_filter = hisq_params.force_filter > 0.0 ? hisq_params.force_filter : 0.0;
}

void qudaHisqForce(int precision, int num_terms, int num_naik_terms, double dt, double **coeff, void **quark_field,
                   const double level2_coeff[6], const double fat7_coeff[6], const void *const w_link,
                   const void *const v_link, const void *const u_link, void *const milc_momentum) {
  static const char where[] = "qudaHisqForce";
  ensure_ctx(where);
  if (num_naik_terms < 0 || num_naik_terms > num_terms) {
    printf("%s: num_naik_terms = %d out of range\n", where, num_naik_terms);
    exit(1);
  }
  std::vector<double> cf(2 * (size_t)(num_terms + num_naik_terms));
  for (int t = 0; t < num_terms + num_naik_terms; t++) {
    cf[2 * t] = coeff[t][0];
    cf[2 * t + 1] = coeff[t][1];
  }
  if (b200ks_hisq_force(S.ctx, num_terms, num_naik_terms, cf.data(), (const void *const *)quark_field, level2_coeff, fat7_coeff, w_link, v_link,
                        u_link, dt, S.force_filter, milc_momentum, precision) < 0)
    die(where);
}

double qudaMomAction(int precision, QudaMILCSiteArg_t *arg) {
  static const char where[] = "qudaMomAction";
  ensure_ctx(where);
  const long nsites = (long)S.latsize[0] * S.latsize[1] * S.latsize[2] * S.latsize[3];
  const char *base = (const char *)(arg->mom ? arg->mom : arg->site);
  const size_t stride = arg->mom ? (size_t)4 * 10 * (precision == 2 ? 8 : 4) : arg->size;
  const size_t off = arg->mom ? 0 : arg->mom_offset;
  char *d = nullptr;
  double *d_out = nullptr, *d_part = nullptr;
  unsigned *d_cnt = nullptr;
  const int grid = (int)((nsites + b200ks::kBlock - 1) / b200ks::kBlock);
  bool ok = cudaMalloc(&d, (size_t)nsites * stride) == cudaSuccess && cudaMalloc(&d_out, sizeof(double)) == cudaSuccess &&
            cudaMalloc(&d_part, sizeof(double) * grid) == cudaSuccess && cudaMalloc(&d_cnt, sizeof(unsigned)) == cudaSuccess;
  double h = 0;
  if (ok) {
    cudaMemset(d_cnt, 0, sizeof(unsigned));
    ok = cudaMemcpy(d, base, (size_t)nsites * stride, cudaMemcpyHostToDevice) == cudaSuccess;
    b200ks::ReduceWs ws{d_part, d_cnt};
    mom_action_kernel<<<grid, b200ks::kBlock>>>(d, off, stride, nsites, precision, ws, d_out);
    ok = ok && cudaMemcpy(&h, d_out, sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  cudaFree(d); cudaFree(d_out); cudaFree(d_part); cudaFree(d_cnt);
  if (!ok) {
    printf("%s: CUDA failure: %s\n", where, cudaGetErrorString(cudaGetLastError()));
    exit(1);
  }
  return h;
}

}  // extern "C"
