// meson.inl -- host side of the meson tie-ups (meson.cuh); included at the end of b200ks.cu.
//
// b200ks_meson_mom[_dev]: corr[t][p] = sum_{x in time slice t} sign_spin(x - r0) <antiquark(x)|quark(x)> ftfact_p(x - r0),
// everything ks_meson_cont_mom (generic_ks/ks_meson_mom.c:160-437) does between the sink operator and norm_v for one
// sink spin-taste assignment.  The caller (csrc_milc/milc_shim.c ks_meson_cont_mom_gpu, api.py) applies the
// correlator phase / factor and accumulates into prop[m][t] as norm_v and the loop behind it do (:100-131, :405-419).

namespace {

constexpr double kMilcPi = 3.14159265358979323846;   // include/complex.h PI

// per-direction factors of ftfact (ff(), ks_meson_mom.c:137-157, arguments as at :276-278): tab[p][coordinate]
int meson_tables(const Geom &g, const int r0[4], int nmom, const int *mom, const char *mpar, std::vector<double2> &tab) {
  const int gsum = g.G[0] + g.G[1] + g.G[2];
  tab.assign((size_t)nmom * gsum, make_double2(0.0, 0.0));
  for (int p = 0; p < nmom; p++) {
    int off = 0;
    for (int d = 0; d < 3; d++) {
      const double fact = 2.0 * kMilcPi / (1.0 * g.G[d]);
      const int e = mpar[3 * p + d];
      if (e != B200KS_EVEN && e != B200KS_ODD && e != B200KS_EVENANDODD)
        return fail(B200KS_EINVAL, "meson tie-up: momentum-component parity must be EVEN (2), ODD (1) or EVENANDODD (3)");
      for (int x = 0; x < g.G[d]; x++) {
        const double theta = fact * (x - r0[d]) * mom[3 * p + d];
        double2 f;
        if (e == B200KS_EVEN) f = make_double2(cos(theta), 0.0);
        else if (e == B200KS_ODD) f = make_double2(0.0, sin(theta));
        else f = make_double2(cos(theta), sin(theta));
        tab[(size_t)p * gsum + off + x] = f;
      }
      off += g.G[d];
    }
  }
  return 0;
}

// one context (plain, or a member of a multi-GPU context): this context's share is ADDED into corr[G[3]][nmom][2]
int meson_local(b200ks_ctx *c, const DevVec &anti, const DevVec &quark, int spin, const int r0[4], int nmom, const int *mom,
                const char *mpar, double *corr) {
  const Geom &g = c->g;
  if (anti.prec != 2 || quark.prec != 2) return fail(B200KS_EINVAL, "meson tie-up: double-precision device vectors only");
  std::vector<double2> tab;
  CHK(meson_tables(g, r0, nmom, mom, mpar, tab));
  const int Lt = g.L[3], slice_h = g.Vh / Lt, nchunk = (slice_h + kMesonSites - 1) / kMesonSites;
  const size_t tab_b = sizeof(double2) * tab.size();
  const size_t part_b = sizeof(double2) * (size_t)Lt * 2 * nchunk * nmom, out_b = sizeof(double2) * (size_t)Lt * nmom;
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  char *buf = nullptr;
  CHK(stage_get(c, up(tab_b) + up(part_b) + up(out_b), (void **)&buf));
  MesonArg a;
  for (int p = 0; p < 2; p++) {
    a.anti[p] = (const double2 *)anti.p[p];
    a.quark[p] = (const double2 *)quark.p[p];
  }
  a.tab = (const double2 *)buf;
  a.partial = (double2 *)(buf + up(tab_b));
  double2 *d_out = (double2 *)(buf + up(tab_b) + up(part_b));
  a.nmom = nmom;
  a.nchunk = nchunk;
  a.slice_h = slice_h;
  a.spin = spin;
  for (int d = 0; d < 4; d++) a.r0[d] = r0[d];
  a.g = g;
  CHK(h2d(c, buf, tab.data(), tab_b));
  meson_mom_kernel<<<dim3(nchunk, Lt, 2), kBlock, 0, c->stream>>>(a);
  meson_finish_kernel<<<(Lt * nmom + 127) / 128, 128, 0, c->stream>>>(a.partial, Lt, nchunk, nmom, d_out);
  c->launches += 2;
  CHK(check_launch("meson_mom_kernel"));
  std::vector<double2> out((size_t)Lt * nmom);
  CHK(d2h(c, out.data(), d_out, out_b));
  for (int t = 0; t < Lt; t++)
    for (int p = 0; p < nmom; p++) {
      double *o = corr + ((size_t)(t + g.origin[3]) * nmom + p) * 2;
      o[0] += out[(size_t)t * nmom + p].x;
      o[1] += out[(size_t)t * nmom + p].y;
    }
  return 0;
}

int meson_check(b200ks_ctx *c, int spin, const int *r0, int nmom, const int *mom, const char *mpar, const double *corr) {
  if (!c || !r0 || !corr || nmom < 1 || !mom || !mpar) return fail(B200KS_EINVAL, "meson tie-up: bad argument");
  if (nmom > kMesonMaxMom) return fail(B200KS_EINVAL, "meson tie-up: at most " + std::to_string(kMesonMaxMom) + " momenta per call");
  if (spin < -1 || spin > 15) return fail(B200KS_EINVAL, "meson tie-up: spin must be -1 (none) or the gamma bits 0..15 of a local operator");
  return 0;
}

// members' shares, each into its own buffer, added in member order (deterministic)
template <typename F>
int meson_all(b200ks_ctx *c, int nmom, double *corr, F per_context) {
  const size_t n = (size_t)c->global[3] * nmom * 2;
  std::fill(corr, corr + n, 0.0);
  if (c->sub.empty()) {
    CU(cudaSetDevice(c->device));
    return per_context(c, corr);
  }
  std::vector<std::vector<double>> share(c->sub.size(), std::vector<double>(n, 0.0));
  CHK(run_all(c, [&](b200ks_ctx *m, int r) -> int {
    CU(cudaSetDevice(m->device));
    return per_context(m, share[r].data());
  }));
  for (size_t r = 0; r < share.size(); r++)
    for (size_t k = 0; k < n; k++) corr[k] += share[r][k];
  return 0;
}

}  // namespace

extern "C" int b200ks_meson_mom_dev(b200ks_ctx *c, int vantiquark, int vquark, int spin, const int *r0, int nmom, const int *mom,
                                    const char *mom_parity, double *corr) {
  CHK(meson_check(c, spin, r0, nmom, mom, mom_parity, corr));
  return meson_all(c, nmom, corr, [&](b200ks_ctx *m, double *out) -> int {
    DevVec *a = uvec(m, vantiquark), *q = uvec(m, vquark);
    if (!a || !q) return B200KS_EINVAL;
    return meson_local(m, *a, *q, spin, r0, nmom, mom, mom_parity, out);
  });
}

extern "C" int b200ks_meson_mom(b200ks_ctx *c, const void *antiquark, const void *quark, int host_prec, int spin, const int *r0,
                                int nmom, const int *mom, const char *mom_parity, double *corr) {
  CHK(meson_check(c, spin, r0, nmom, mom, mom_parity, corr));
  if (!antiquark || !quark) return fail(B200KS_EINVAL, "b200ks_meson_mom: null field");
  return meson_all(c, nmom, corr, [&](b200ks_ctx *m, double *out) -> int {
    DevVec *a = nullptr, *q = nullptr;
    CHK(pool_get(m, 2, kBlockPool + 4 * kMaxRhs, &a));       // (the resident UML sequence's upload slots)
    CHK(pool_get(m, 2, kBlockPool + 4 * kMaxRhs + 1, &q));
    CHK(upload(m, *a, antiquark, B200KS_EVENANDODD, host_prec, false));
    if (quark == antiquark) q = a;
    else CHK(upload(m, *q, quark, B200KS_EVENANDODD, host_prec, false));
    return meson_local(m, *a, *q, spin, r0, nmom, mom, mom_parity, out);
  });
}
