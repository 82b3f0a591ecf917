// eigcg.inl -- host side of the eigCG / incremental eigCG solver (included at the end of b200ks.cu, which
// owns the context).  Replaces ks_eigCG_parity, ks_inc_eigCG_parity and calc_eigenpairs
// (generic_ks/inc_eigcg.c:377-850, 851-950, 282-300), called by mat_invert_uml_field when MILC is built with
// EIGMODE = EIGCG (generic_ks/mat_invert.c:361-363).
//
// The CG itself is the context's pure-double single-mass solver iteration for iteration (same kernels, same
// FEWSUMS arithmetic as b200ks_congrad): the eigCG part only OBSERVES it -- the coefficients a_k, b_k build
// the Lanczos matrix T, the residual before each update is normalised into the search window.  The host is in
// the loop once per iteration (it needs a_k, b_k and runs the window restarts), which costs one stream
// synchronisation per ~0.5 ms iteration at 32^3 x 64.
#include "eigcg.cuh"

struct EigCGState {
  int m = 0, nvecs = 0, ncurr = 0, nmax = 0;
  int pbit = -1;                        // parity bit the accumulated vectors live on (-1: none yet)
  std::vector<double2 *> slot;          // nmax + m + 1 parity-half vectors: [accumulated | window | spare]
  std::vector<double2 *> tmp;           // rotation outputs
  double2 **d_ptr = nullptr;            // device copy of slot[] (+ tmp[] behind it)
  double *d_part = nullptr, *d_dots = nullptr;
  double2 *d_coef = nullptr;
  double2 *ttt2 = nullptr;
  int nchunks = 0;
  size_t coef_cap = 0;
  int nslot = 0, ntmp = 0;              // as allocated (bookkeeping of the releases)
  std::vector<dense::cd> H;             // nmax x nmax, row-major, upper triangle = -U^+ D^2 U
  std::vector<double> val;
  size_t vbytes = 0;
};

static void eigcg_release(b200ks_ctx *c) {
  EigCGState *e = (EigCGState *)c->eigcg;
  if (!e) return;
  cudaStreamSynchronize(c->stream);
  for (auto p : e->slot) dev_free(c, p, e->vbytes);
  for (auto p : e->tmp) dev_free(c, p, e->vbytes);
  dev_free(c, e->ttt2, e->vbytes);
  dev_free(c, e->d_ptr, sizeof(double2 *) * (size_t)(e->nslot + e->ntmp));
  dev_free(c, e->d_part, sizeof(double) * 4 * (size_t)e->nslot * e->nchunks);
  dev_free(c, e->d_dots, sizeof(double) * 4 * (size_t)e->nslot);
  dev_free(c, e->d_coef, sizeof(double2) * e->coef_cap);
  delete e;
  c->eigcg = nullptr;
}

extern "C" int b200ks_eigcg_init(b200ks_ctx *c, int m, int nvecs, int nvecs_max) {
  if (!c) return fail(B200KS_EINVAL, "null context");
  if (!c->sub.empty() || c->comm.active) return fail(B200KS_ESTATE, "eigCG: single-GPU contexts only");
  if (nvecs < 0 || m < 2 || 2 * nvecs >= m || nvecs_max < nvecs)
    return fail(B200KS_EINVAL, "b200ks_eigcg_init: need 2*Nvecs < m and Nvecs <= Nvecs_max (generic_ks/inc_eigcg.c:377-400)");
  CU(cudaSetDevice(c->device));
  eigcg_release(c);
  EigCGState *e = new EigCGState;
  c->eigcg = e;
  e->m = m; e->nvecs = nvecs; e->nmax = nvecs_max;
  e->vbytes = (size_t)3 * c->g.stride * sizeof(double2);
  const int nslot = nvecs_max + m + 1, ntmp = std::max(2 * nvecs, kRotOut);
  e->nslot = nslot;
  e->ntmp = ntmp;
  e->nchunks = std::min(nblocks(c->g.Vh), 1184);
  e->coef_cap = (size_t)std::max(nslot, nvecs_max) * std::max(ntmp, nvecs_max);
  int r = 0;
  for (int j = 0; j < nslot && r == 0; j++) {
    void *p = nullptr;
    r = dev_alloc(c, &p, e->vbytes);
    if (r == 0) { e->slot.push_back((double2 *)p); cudaMemsetAsync(p, 0, e->vbytes, c->stream); }
  }
  for (int j = 0; j < ntmp && r == 0; j++) {
    void *p = nullptr;
    r = dev_alloc(c, &p, e->vbytes);
    if (r == 0) e->tmp.push_back((double2 *)p);
  }
  void *p = nullptr;
  if (r == 0) { r = dev_alloc(c, &p, e->vbytes); e->ttt2 = (double2 *)p; }
  if (r == 0) { r = dev_alloc(c, &p, sizeof(double2 *) * (nslot + ntmp)); e->d_ptr = (double2 **)p; }
  if (r == 0) { r = dev_alloc(c, &p, sizeof(double) * 4 * (size_t)nslot * e->nchunks); e->d_part = (double *)p; }
  if (r == 0) { r = dev_alloc(c, &p, sizeof(double) * 4 * (size_t)nslot); e->d_dots = (double *)p; }
  if (r == 0) { r = dev_alloc(c, &p, sizeof(double2) * e->coef_cap); e->d_coef = (double2 *)p; }
  if (r < 0) {
    eigcg_release(c);
    return r;
  }
  e->H.assign((size_t)nvecs_max * nvecs_max, dense::cd(0, 0));
  e->val.assign(nvecs_max + m, 0.0);
  return 0;
}

static int eig_ptrs_push(b200ks_ctx *c, EigCGState *e) {
  std::vector<double2 *> all(e->slot);
  all.insert(all.end(), e->tmp.begin(), e->tmp.end());
  CU(cudaMemcpyAsync(e->d_ptr, all.data(), sizeof(double2 *) * all.size(), cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));   // `all` is a local
  return 0;
}

// out[j] = { <slot[first+j] | a>, <slot[first+j] | b> }, j < n
static int eig_dots(b200ks_ctx *c, EigCGState *e, int first, int n, const double2 *a, const double2 *b, std::vector<dense::cd> &da,
                    std::vector<dense::cd> *db) {
  da.assign(n, dense::cd(0, 0));
  if (db) db->assign(n, dense::cd(0, 0));
  if (n == 0) return 0;
  const int nv = c->g.Vh, per = (nv + e->nchunks - 1) / e->nchunks;
  eig_dot_kernel<<<e->nchunks, kBlock, 0, c->stream>>>((const double2 *const *)(e->d_ptr + first), n, a, b, c->g.stride, nv, per, e->d_part);
  eig_dotsum_kernel<<<(n + 3) / 4, 128, 0, c->stream>>>(e->d_part, e->nchunks, n, e->d_dots);
  c->launches += 2;
  std::vector<double> h(4 * (size_t)n);
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(h.data(), e->d_dots, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; j++) {
    da[j] = dense::cd(h[4 * j], h[4 * j + 1]);
    if (db) (*db)[j] = dense::cd(h[4 * j + 2], h[4 * j + 3]);
  }
  return check_launch("eig_dot_kernel");
}

// dst += sum_j coef[j] slot[first + j]
static int eig_combine(b200ks_ctx *c, EigCGState *e, int first, int n, const std::vector<dense::cd> &coef, double2 *dst) {
  if (n == 0) return 0;
  std::vector<double2> h(n);
  for (int j = 0; j < n; j++) h[j] = make_double2(coef[j].real(), coef[j].imag());
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(e->d_coef, h.data(), sizeof(double2) * n, cudaMemcpyHostToDevice));
  eig_axpy_kernel<<<nblocks(c->g.Vh), kBlock, 0, c->stream>>>((const double2 *const *)(e->d_ptr + first), e->d_coef, n, dst, c->g.stride, c->g.Vh);
  c->launches++;
  return check_launch("eig_axpy_kernel");
}

// slot[first + o] <- sum_jj C[jj][o] slot[first + jj]  (jj < nin, o < nout <= tmp.size()), through the tmp vectors
static int eig_rotate(b200ks_ctx *c, EigCGState *e, int first, int nin, int nout, const std::vector<dense::cd> &C) {
  if (nout > (int)e->tmp.size() || (size_t)nin * nout > e->coef_cap) return fail(B200KS_ESTATE, "eig_rotate: workspace too small");
  std::vector<double2> h((size_t)nin * nout);
  for (size_t k = 0; k < h.size(); k++) h[k] = make_double2(C[k].real(), C[k].imag());
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(e->d_coef, h.data(), sizeof(double2) * h.size(), cudaMemcpyHostToDevice));
  dim3 grid(nblocks(c->g.Vh), (nout + kRotOut - 1) / kRotOut);
  eig_rotate_kernel<<<grid, kBlock, 0, c->stream>>>((const double2 *const *)(e->d_ptr + first), nin, e->d_ptr + e->slot.size(), nout, e->d_coef,
                                                    c->g.stride, c->g.Vh);
  c->launches++;
  for (int o = 0; o < nout; o++) std::swap(e->slot[first + o], e->tmp[o]);   // results become the first nout vectors
  CHK(eig_ptrs_push(c, e));
  return check_launch("eig_rotate_kernel");
}

static void scale_copy(b200ks_ctx *c, double2 *out, double a, const double2 *x, double b, const double2 *y) {
  LAUNCH(c, (axpby_kernel<double>), nblocks(c->g.Vh), out, a, x, b, y, c->g.stride, c->g.Vh);
}

// ks_eigCG_parity (inc_eigcg.c:377-850) on window slots [w0, w0 + m]; Ritz values of -D^2 into val[w0 ..]
static int eigcg_solve(b200ks_ctx *c, EigCGState *e, int w0, int nvecs, const DevVec &b, DevVec &x, double mass,
                       const b200ks_invert_args &args, b200ks_invert_result &res) {
  using dense::cd;
  const int pb = parity_bit(args.parity), ob = pb ^ 1;
  const Geom &g = c->g;
  const int grid = nblocks(g.Vh);
  const int m = e->m, niter = args.max_iter, max_restarts = args.nrestart;
  const double rsqmin = args.resid * args.resid;
  const double msq_x4 = 4.0 * mass * mass;
  const int max_cg = max_restarts * niter;
  if (args.relresid != 0) return fail(B200KS_EINVAL, "eigCG: Fermilab relative residual not supported");
  res = b200ks_invert_result();
  res.converged = 1;
  res.size_relr = 1.0;
  double source_norm = 0;
  CHK(norm2(c, b, pb, &source_norm));
  if (source_norm == 0.0) {  // inc_eigcg.c:461-475
    CHK(zero_half(c, x, pb));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
  }
  DevVec *ttt, *p, *r;
  CHK(pool_get(c, 2, 2, &ttt));
  CHK(pool_get(c, 2, 3, &p));
  CHK(pool_get(c, 2, 4, &r));
  CgState &h = *c->h_state;
  memset(&h, 0, sizeof(h));
  h.source_norm = source_norm;
  h.rsqmin = rsqmin;
  h.size_relr = 1.0;
  h.niter = niter;
  h.half_volume = 0.5 * (double)c->global[0] * c->global[1] * c->global[2] * c->global[3];

  std::vector<cd> T((size_t)m * m, cd(0, 0));   // row-major, upper triangle
  auto Tat = [&](int i, int j) -> cd & { return T[(size_t)i * m + j]; };
  int k = -1;
  double a = 1.0, bcoef = 0.0;
  int iteration = 0, nrestart = 0;
  CU(cudaEventRecord(c->ev0, c->stream));
  for (;;) {
    {   // (re)start from the true residual, inc_eigcg.c:520-576
      Epi e0, e1;
      CHK(dslash_T<double>(c, x, *ttt, ob, e0));
      e1.kind = 1; e1.s = -msq_x4; e1.w = &x;
      CHK(dslash_T<double>(c, *ttt, *ttt, pb, e1));
      LAUNCH(c, (cg_restart_kernel<double, false>), grid, (const double2 *)b.p[pb], (const double2 *)ttt->p[pb], (const double2 *)x.p[pb],
             (double2 *)r->p[pb], (double2 *)p->p[pb], g.stride, g.Vh, c->ws, c->d_scal);
      CU(cudaMemcpyAsync(c->h_scal, c->d_scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CU(cudaStreamSynchronize(c->stream));
      CHK(check_launch("eigcg restart"));
      const double rsq = c->h_scal[0];
      res.final_rsq = rsq / source_norm;
      iteration++;
      if (iteration >= max_cg || nrestart >= max_restarts || (rsqmin <= 0 || rsqmin > res.final_rsq)) break;
      if (nvecs > 0) {
        a = 1.0; bcoef = 0.0; k = -1;
        std::fill(T.begin(), T.end(), cd(0, 0));
      }
      nrestart++;
      h.rsq = rsq;
      h.upd[0] = rsq;
      h.upd[1] = 0;
      h.iter = iteration;
      h.stop = 0;
      CHK(state_push(c));
    }
    double rsq_prev = h.rsq, upd_prev = h.upd[0];
    for (;;) {   // one CG iteration per pass, the host in the loop
      Epi e0, e1;
      CHK(dslash_T<double>(c, *p, *ttt, ob, e0));
      e1.kind = 2; e1.s = -msq_x4; e1.w = p; e1.r = r; e1.red = c->d_state->red;
      CHK(dslash_T<double>(c, *ttt, *ttt, pb, e1));
      if (nvecs > 0) {
        if (k == m - 1) {   // window full: compress it to 2 Nvecs Ritz vectors, inc_eigcg.c:607-690
          const int n2 = 2 * nvecs;
          std::vector<double> w1, w2;
          std::vector<cd> Z1, Z2, T1((size_t)(m - 1) * (m - 1));
          dense::heev(m, T, w1, Z1);
          for (int i = 0; i < m - 1; i++)
            for (int j = 0; j < m - 1; j++) T1[(size_t)i * (m - 1) + j] = Tat(i, j);
          dense::heev(m - 1, T1, w2, Z2);
          std::vector<cd> Y((size_t)m * n2, cd(0, 0));
          for (int i = 0; i < m; i++)
            for (int j = 0; j < nvecs; j++) Y[(size_t)i * n2 + j] = Z1[(size_t)i * m + j];
          for (int i = 0; i < m - 1; i++)
            for (int j = 0; j < nvecs; j++) Y[(size_t)i * n2 + nvecs + j] = Z2[(size_t)i * (m - 1) + j];
          dense::orthonormalize(m, n2, Y);
          // Ts = Q^+ T Q (T Hermitian from its upper triangle)
          std::vector<cd> TQ((size_t)m * n2, cd(0, 0)), Ts((size_t)n2 * n2, cd(0, 0));
          for (int i = 0; i < m; i++)
            for (int l = 0; l < m; l++) {
              const cd t = (l >= i) ? Tat(i, l) : std::conj(Tat(l, i));
              if (t == cd(0, 0)) continue;
              for (int j = 0; j < n2; j++) TQ[(size_t)i * n2 + j] += t * Y[(size_t)l * n2 + j];
            }
          for (int i = 0; i < n2; i++)
            for (int j = 0; j < n2; j++) {
              cd s = 0;
              for (int l = 0; l < m; l++) s += std::conj(Y[(size_t)l * n2 + i]) * TQ[(size_t)l * n2 + j];
              Ts[(size_t)i * n2 + j] = s;
            }
          std::vector<double> ev;
          std::vector<cd> Z;
          dense::heev(n2, Ts, ev, Z);
          std::vector<cd> QZ((size_t)m * n2, cd(0, 0));
          for (int i = 0; i < m; i++)
            for (int j = 0; j < n2; j++) {
              cd s = 0;
              for (int l = 0; l < n2; l++) s += Y[(size_t)i * n2 + l] * Z[(size_t)l * n2 + j];
              QZ[(size_t)i * n2 + j] = s;
            }
          CHK(eig_rotate(c, e, w0, m, n2, QZ));
          std::fill(T.begin(), T.end(), cd(0, 0));
          for (int j = 0; j < n2; j++) Tat(j, j) = ev[j];
          k = n2 - 1;
          // ttt2 <- ttt2 - ttt ; T_{j,k+1} = <V_j | ttt2> / sqrt(rsq)
          scale_copy(c, e->ttt2, 1.0, e->ttt2, -1.0, (const double2 *)ttt->p[pb]);
          std::vector<cd> tau;
          CHK(eig_dots(c, e, w0, n2, e->ttt2, e->ttt2, tau, nullptr));
          for (int j = 0; j < n2; j++) Tat(j, k + 1) = tau[j] / std::sqrt(rsq_prev);
        } else if (k >= 0) {
          Tat(k, k + 1) = -std::sqrt(bcoef) / a;
        }
        k++;
        scale_copy(c, e->slot[w0 + k], 1.0 / std::sqrt(rsq_prev), (const double2 *)r->p[pb], 0.0, nullptr);   // V_k = r / |r|
        Tat(k, k) = bcoef / a;
      }
      const int fuse = 1 | 8;
      LAUNCH(c, (cg_update_kernel<double, false>), grid, (double2 *)x.p[pb], (double2 *)r->p[pb], (double2 *)p->p[pb],
             (const double2 *)ttt->p[pb], g.stride, g.Vh, c->d_state, c->ws, fuse);
      finish_update(c, grid, c->d_state, fuse);
      CHK(state_pull(c));
      CHK(check_launch("eigcg iterate"));
      // the coefficients the kernels have just used (blas.cuh cg_update_kernel)
      a = -rsq_prev / h.red[0];
      bcoef = h.rsq / upd_prev;
      iteration = h.iter;
      res.size_r = h.size_r;
      res.final_iters = iteration;
      res.final_restart = nrestart;
      if (nvecs > 0) {
        Tat(k, k) += 1.0 / a;
        if (k == m - 1) scale_copy(c, e->ttt2, bcoef, (const double2 *)ttt->p[pb], 0.0, nullptr);   // ttt2 = b ttt
      }
      rsq_prev = h.rsq;
      upd_prev = h.upd[0];
      if (h.stop) break;   // restart interval reached or recursive residual under the target
    }
  }
  if (nvecs > 0 && k >= 0) {   // final Rayleigh-Ritz on the k+1 window vectors, inc_eigcg.c:795-808
    const int n = k + 1, nout = std::min(nvecs, n);
    std::vector<cd> Tn((size_t)n * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) Tn[(size_t)i * n + j] = Tat(i, j);
    std::vector<double> w;
    std::vector<cd> Z, C((size_t)n * nout);
    dense::heev(n, Tn, w, Z);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < nout; j++) C[(size_t)i * nout + j] = Z[(size_t)i * n + j];
    CHK(eig_rotate(c, e, w0, n, nout, C));
    for (int j = 0; j < nout; j++) e->val[w0 + j] = w[j] - msq_x4;
  }
  CU(cudaEventRecord(c->ev1, c->stream));
  CU(cudaEventSynchronize(c->ev1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  res.device_seconds = ms * 1e-3;
  res.final_iters = iteration;
  res.final_restart = nrestart;
  res.converged = (nrestart == max_restarts || iteration == max_cg) ? 0 : 1;
  return iteration;
}

// ks_inc_eigCG_parity, inc_eigcg.c:851-950
static int inc_eigcg_any(b200ks_ctx *c, const DevVec &b, DevVec &x, double mass, const b200ks_invert_args &args,
                         b200ks_invert_result &res) {
  using dense::cd;
  EigCGState *e = (EigCGState *)c->eigcg;
  if (!e) return fail(B200KS_ESTATE, "eigCG: b200ks_eigcg_init first");
  CHK(links_ensure(c, 2));
  const int pb = parity_bit(args.parity), ob = pb ^ 1;
  if (e->ncurr > 0 && e->pbit != pb) return fail(B200KS_EINVAL, "eigCG: the accumulated vectors live on the other parity");
  e->pbit = pb;
  CHK(eig_ptrs_push(c, e));
  const double msq_x4 = 4.0 * mass * mass;
  const int nc = e->ncurr, nmax = e->nmax;
  DevVec *ttt;
  CHK(pool_get(c, 2, 2, &ttt));
  if (nc == 0) {
    std::fill(e->H.begin(), e->H.end(), cd(0, 0));
  } else {   // initCG, inc_eigcg.c:47-121: x += U (H + 4 m^2)^-1 U^+ (b - A x)
    Epi e0, e1;
    CHK(dslash_T<double>(c, x, *ttt, ob, e0));
    e1.kind = 1; e1.s = -msq_x4; e1.w = &x;
    CHK(dslash_T<double>(c, *ttt, *ttt, pb, e1));
    scale_copy(c, (double2 *)ttt->p[pb], 1.0, (const double2 *)ttt->p[pb], 1.0, (const double2 *)b.p[pb]);   // resid = b + ttt
    std::vector<cd> cvec;
    CHK(eig_dots(c, e, 0, nc, (const double2 *)ttt->p[pb], (const double2 *)ttt->p[pb], cvec, nullptr));
    std::vector<cd> H2((size_t)nc * nc);
    for (int i = 0; i < nc; i++)
      for (int j = 0; j < nc; j++) H2[(size_t)i * nc + j] = e->H[(size_t)i * nmax + j] + (i == j ? cd(msq_x4, 0) : cd(0, 0));
    if (!dense::posv(nc, H2, cvec)) return fail(B200KS_ESTATE, "eigCG: H + 4 m^2 is not positive definite");
    CHK(eig_combine(c, e, 0, nc, cvec, (double2 *)x.p[pb]));
  }
  const int it = eigcg_solve(c, e, nc, e->nvecs, b, x, mass, args, res);
  if (it < 0) return it;
  if (e->nvecs > 0) {
    // orthogonalize, inc_eigcg.c:123-213 (classical Gram-Schmidt against everything before, twice: one pass of
    // n dot products per new vector instead of the reference's n sequential passes)
    int j = nc, add = e->nvecs, n = nc + add;
    while (j < n) {
      for (int pass = 0; pass < 2; pass++) {
        std::vector<cd> d;
        CHK(eig_dots(c, e, 0, j, e->slot[j], e->slot[j], d, nullptr));
        for (auto &z : d) z = -z;
        CHK(eig_combine(c, e, 0, j, d, e->slot[j]));
      }
      std::vector<cd> nn;
      CHK(eig_dots(c, e, j, 1, e->slot[j], e->slot[j], nn, nullptr));
      const double norm = std::sqrt(nn[0].real());
      if (norm < 1e-15) {   // ORTHO_EPS, include/imp_ferm_links.h:409
        add--;
        n--;
        double2 *dead = e->slot[j];
        for (int q = j; q < n; q++) e->slot[q] = e->slot[q + 1];
        e->slot[n] = dead;
        CHK(eig_ptrs_push(c, e));
      } else {
        scale_copy(c, e->slot[j], 1.0 / norm, e->slot[j], 0.0, nullptr);
        j++;
      }
    }
    // extend_H, inc_eigcg.c:215-262: H_{k,j} = -<U_k | D^2 U_j> for the new columns
    DevVec *tv;
    CHK(pool_get(c, 2, 3, &tv));
    for (int q = nc; q < nc + add; q++) {
      CU(cudaMemcpyAsync(tv->p[pb], e->slot[q], e->vbytes, cudaMemcpyDeviceToDevice, c->stream));
      Epi e0;
      CHK(dslash_T<double>(c, *tv, *ttt, ob, e0));
      CHK(dslash_T<double>(c, *ttt, *ttt, pb, e0));
      std::vector<cd> d;
      CHK(eig_dots(c, e, 0, nc + add, (const double2 *)ttt->p[pb], (const double2 *)ttt->p[pb], d, nullptr));
      for (int kk = 0; kk < nc + add; kk++) e->H[(size_t)kk * nmax + q] = -d[kk];
    }
    e->ncurr = nc + add;
    e->nvecs = std::min(nmax - e->ncurr, e->nvecs);
  }
  return it;
}

extern "C" int b200ks_inc_eigcg_dev(b200ks_ctx *c, int vsrc, int vdest, double mass, const b200ks_invert_args *args,
                                    b200ks_invert_result *res) {
  if (c && !c->sub.empty()) return fail(B200KS_ESTATE, "eigCG: single-GPU contexts only");
  DevVec *b = uvec(c, vsrc), *x = uvec(c, vdest);
  if (!b || !x || !res) return fail(B200KS_EINVAL, "b200ks_inc_eigcg_dev: bad argument");
  if (b == x) return fail(B200KS_EINVAL, "source and solution must be different fields");
  CHK(check_args(args));
  CU(cudaSetDevice(c->device));
  return inc_eigcg_any(c, *b, *x, mass, *args, *res);
}

extern "C" int b200ks_inc_eigcg(b200ks_ctx *c, const void *src, void *dest, double mass, const b200ks_invert_args *args,
                                b200ks_invert_result *res, int host_prec) {
  if (!c || !src || !dest || !res) return fail(B200KS_EINVAL, "b200ks_inc_eigcg: null argument");
  if (!c->sub.empty()) return fail(B200KS_ESTATE, "eigCG: single-GPU contexts only");
  CHK(check_args(args));
  CU(cudaSetDevice(c->device));
  DevVec *b = nullptr, *x = nullptr;
  CHK(pool_get(c, 2, 0, &b));
  CHK(pool_get(c, 2, 1, &x));
  int it = with_verified_links(c, [&]() -> int {
    CHK(upload(c, *b, src, args->parity, host_prec, false));
    CHK(upload(c, *x, dest, args->parity, host_prec, false));
    return inc_eigcg_any(c, *b, *x, mass, *args, *res);
  });
  if (it < 0) return it;
  CHK(download(c, *x, dest, args->parity, host_prec));
  return it;
}

// calc_eigenpairs (inc_eigcg.c:282-300 + RayleighRitz :264-280): rotates the accumulated vectors into the Ritz
// basis of H; eigval (may be NULL) receives the Ritz values of -D^2, ascending.  Returns their number.
extern "C" int b200ks_eigcg_pairs(b200ks_ctx *c, double *eigval, int nmax_out) {
  using dense::cd;
  if (!c || !c->eigcg) return fail(B200KS_ESTATE, "eigCG: b200ks_eigcg_init first");
  EigCGState *e = (EigCGState *)c->eigcg;
  CU(cudaSetDevice(c->device));
  const int n = e->ncurr, nmax = e->nmax;
  if (n == 0) return 0;
  std::vector<cd> Hn((size_t)n * n), Z;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) Hn[(size_t)i * n + j] = e->H[(size_t)i * nmax + j];
  std::vector<double> w;
  dense::heev(n, Hn, w, Z);
  CHK(eig_ptrs_push(c, e));
  // V <- V Z: every output depends on every input, so the outputs go to n fresh vectors (ntmp at a time through the
  // rotation kernel) that replace the old ones at the end -- twice the set in HBM for a moment (n = 500 at
  // 32^3 x 64: 2 x 25 GB of 180)
  std::vector<double2 *> fresh;
  auto drop = [&]() { for (auto q : fresh) dev_free(c, q, e->vbytes); };
  for (int j = 0; j < n; j++) {
    void *q = nullptr;
    const int r = dev_alloc(c, &q, e->vbytes);
    if (r < 0) { drop(); return r; }
    fresh.push_back((double2 *)q);
  }
  const int blk = e->ntmp;
  for (int j0 = 0; j0 < n; j0 += blk) {
    const int nb = std::min(blk, n - j0);
    std::vector<double2> hc((size_t)n * nb);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < nb; j++) hc[(size_t)i * nb + j] = make_double2(Z[(size_t)i * n + j0 + j].real(), Z[(size_t)i * n + j0 + j].imag());
    if (hc.size() > e->coef_cap) { drop(); return fail(B200KS_ESTATE, "b200ks_eigcg_pairs: workspace too small"); }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(e->d_coef, hc.data(), sizeof(double2) * hc.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_ptr + e->nslot, fresh.data() + j0, sizeof(double2 *) * nb, cudaMemcpyHostToDevice));
    dim3 grid(nblocks(c->g.Vh), (nb + kRotOut - 1) / kRotOut);
    eig_rotate_kernel<<<grid, kBlock, 0, c->stream>>>((const double2 *const *)e->d_ptr, n, e->d_ptr + e->nslot, nb, e->d_coef, c->g.stride, c->g.Vh);
    c->launches++;
  }
  CU(cudaStreamSynchronize(c->stream));
  for (int j = 0; j < n; j++) std::swap(e->slot[j], fresh[j]);
  drop();
  CHK(eig_ptrs_push(c, e));
  CHK(check_launch("eig_rotate_kernel"));
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) e->H[(size_t)i * nmax + j] = (i == j) ? cd(w[i], 0) : cd(0, 0);
  for (int j = 0; j < n; j++) {
    e->val[j] = w[j];
    if (eigval && j < nmax_out) eigval[j] = w[j];
  }
  return n;
}

extern "C" int b200ks_eigcg_count(b200ks_ctx *c) { return (c && c->eigcg) ? ((EigCGState *)c->eigcg)->ncurr : 0; }

// eigenvector j of the accumulated set into a MILC-order host field (the eigCG parity; the other parity untouched)
extern "C" int b200ks_eigcg_vec_download(b200ks_ctx *c, int j, void *host, int host_prec) {
  if (!c || !c->eigcg || !host) return fail(B200KS_EINVAL, "b200ks_eigcg_vec_download: bad argument");
  EigCGState *e = (EigCGState *)c->eigcg;
  if (j < 0 || j >= e->ncurr) return fail(B200KS_EINVAL, "b200ks_eigcg_vec_download: no such vector");
  CU(cudaSetDevice(c->device));
  DevVec *t = nullptr;
  CHK(pool_get(c, 2, 3, &t));
  CU(cudaMemcpyAsync(t->p[e->pbit], e->slot[j], e->vbytes, cudaMemcpyDeviceToDevice, c->stream));
  return download(c, *t, host, e->pbit ? B200KS_ODD : B200KS_EVEN, host_prec);
}

// H = -U^+ D^2 U as accumulated (row-major nmax x nmax complex, upper triangle meaningful), for the host's eigcg_params
extern "C" int b200ks_eigcg_hmatrix(b200ks_ctx *c, double *H_out) {
  if (!c || !c->eigcg || !H_out) return fail(B200KS_EINVAL, "b200ks_eigcg_hmatrix: bad argument");
  EigCGState *e = (EigCGState *)c->eigcg;
  memcpy(H_out, e->H.data(), sizeof(dense::cd) * e->H.size());
  return e->nmax;
}
