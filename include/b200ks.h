/* include/b200ks.h -- C ABI of the B200-native improved-staggered (HISQ/asqtad) solver.
 *
 * Plain C, plain pointers and sizes; no MILC types, no torch types.  This is the
 * drop-in boundary for ONE hot path of milc-qcd/milc_qcd: the fat+Naik Dirac stencil
 * inside the single-mass CG and the multi-shift CG.  Each entry point names the
 * reference interface it replaces (paths relative to the MILC tree).
 *
 * Host arrays are MILC's own layout:
 *   site index i = node_index(x,y,z,t): lex = x + nx*(y + ny*(z + nz*t)), even sites
 *   first (lex/2) then odd ((lex+V)/2)          generic/layout_hyper_prime.c:509-520
 *   su3_vector  = 3 complex  (re,im) -> 6 reals  include/milc_datatypes.h:49,56
 *   su3_matrix  = e[row][col] complex -> 18 reals, links stored fat[4*i+dir], dir=X,Y,Z,T
 *                                               generic_ks/dslash_fn.c:453,458
 * host_prec is MILC_PRECISION of the caller: 1 = float arrays, 2 = double arrays.
 *
 * Error convention: functions return 0 on success and a negative B200KS_E* code on
 * failure; b200ks_last_error() gives the message.  The MILC-facing shims turn a failure
 * into MILC's own convention, printf + terminate(1) (generic/com_vanilla.c:203-211).
 * There is NO CPU fallback: without a usable sm_100 device b200ks_create fails.
 *
 * Threading: like the reference (file-static temporaries, generic_ks/dslash_fn.c:26-28)
 * a context is single-threaded and non-reentrant.
 */
#ifndef B200KS_H
#define B200KS_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200KS_VERSION 121 /* 110: block solve, resident sequences, link construction; 111: force filter, Naik epsilons, deflation;
                              120: single-process multi-GPU contexts, b200ks_links_sync, flag-based reductions, eigCG; 121: meson tie-ups */

/* parity codes, include/macros.h:68-70 */
#define B200KS_EVEN 2
#define B200KS_ODD 1
#define B200KS_EVENANDODD 3

/* device arithmetic/storage precision of a solve or a dslash */
#define B200KS_PREC_HALF 0   /* int16 fixed point + fp32 site norm (inner solves only) */
#define B200KS_PREC_SINGLE 1
#define B200KS_PREC_DOUBLE 2

#define B200KS_EINVAL (-1)
#define B200KS_ECUDA (-2)
#define B200KS_ENOMEM (-3)
#define B200KS_ESTATE (-4)
#define B200KS_ECOMM (-5)

#define B200KS_MAX_SHIFTS 32 /* MAX_MMINV_NMASSES, include/imp_ferm_links.h:485 */

typedef struct b200ks_ctx b200ks_ctx;

/* Solver controls = the fields of quark_invert_control the solvers read
 * (include/generic_quark_types.h:167-190) + QudaInvertArgs_t.mixed_precision
 * (generic_ks/d_congrad5_fn_gpu.c:104-111). */
typedef struct {
  int parity;          /* B200KS_EVEN or B200KS_ODD (qic->parity)                    */
  int max_iter;        /* iterations per restart (qic->max)                          */
  int nrestart;        /* max restarts (qic->nrestart)                               */
  double resid;        /* target sqrt(|r|^2/|b|^2), NOT squared (qic->resid)         */
  double relresid;     /* Fermilab relative residual target, 0 = unused (d_congrad5_fn_milc.c:37-56; with
                          mixed_precision != 0 on a partitioned context the solve runs in pure double) */
  int mixed_precision; /* 0 pure double; 1 double solution/true residuals + single-precision
                          Krylov vectors (HALF_MIXED); 2 additionally 16-bit links and search
                          direction in the stencil (MAX_MIXED); both with reliable updates    */
  int check_interval;  /* host convergence poll every n iterations; 0 = default      */
} b200ks_invert_args;

/* Solver outputs = the fields the solvers write back into quark_invert_control
 * (generic_ks/d_congrad5_fn_milc.c:125-133,220-225,350-354,370-381). */
typedef struct {
  double final_rsq;    /* true |r|^2/|b|^2 at exit                                   */
  double final_relrsq; /* Fermilab relative residual at exit                         */
  double size_r;       /* last recursive |r|^2/|b|^2                                 */
  double size_relr;
  int final_iters;
  int final_restart;
  int converged;
  double device_seconds; /* CUDA-event time of the solve proper (no host<->device copies) */
} b200ks_invert_result;

int b200ks_version(void);
const char *b200ks_last_error(void);

/* Number of sm_100 devices visible (0 => nothing can run; callers must fail loudly). */
int b200ks_device_count(void);

/* Single-GPU context for an nx*ny*nz*nt lattice on CUDA device `device`.
 * Replaces initialize_quda()/qudaInit (generic/milc_to_quda_utilities.c:11-44). */
b200ks_ctx *b200ks_create(const int latsize[4], int device);

/* One-rank-per-GPU context: global lattice `latsize` split over `grid[4]` ranks
 * (t first, then z: grid = {1,1,gz,gt}); this process owns the sub-lattice at grid
 * coordinates derived from `rank` (t slowest).  nccl_unique_id is the 128-byte
 * ncclUniqueId shared by all ranks (obtain with b200ks_comm_unique_id on rank 0 and
 * broadcast it by any means).  Host arrays passed to this context are the LOCAL
 * sub-lattice in MILC order, exactly what a MILC MPI rank holds
 * (generic/layout_hyper_prime.c:186-229). */
b200ks_ctx *b200ks_create_dist(const int latsize[4], const int grid[4], int rank, int nranks,
                               const void *nccl_unique_id, int device);
int b200ks_comm_unique_id(void *out128);

/* Single-process multi-GPU context ("single process / N devices, MILC runs as
 * one vanilla rank and never sees the decomposition"; the reference hands QUDA its machine grid in
 * generic/milc_to_quda_utilities.c:13-36, the split itself is generic/layout_hyper_prime.c:186-229).
 * The lattice is split over ngpu devices (devices[0..ngpu-1], NULL = 0..ngpu-1) in t first (up to 4
 * ways), then z; local extents must be even and >= 6.  Every device gets an ordinary partitioned
 * context driven by its own host thread -- the same kernels, peer-to-peer halo pushes and flag-based
 * reductions as the one-rank-per-GPU form, over plain peer access instead of CUDA IPC, no NCCL.  The
 * returned context takes the SAME host arrays as a single-GPU context (MILC's global layout): every
 * device reads and writes its own sub-lattice straight from / to them.  Solves, dslash, the resident
 * sequences and the device-vector interface run decomposed; link construction and the fermion force
 * run on a full-lattice context on devices[0]; deflation is refused.
 * B200KS_NGPU=N in the environment makes b200ks_create (what both MILC-facing shims call) behave as
 * b200ks_create_multi(latsize, N, {device, device+1, ...}). */
b200ks_ctx *b200ks_create_multi(const int latsize[4], int ngpu, const int *devices);
int b200ks_num_gpus(b200ks_ctx *ctx);
/* How this context exchanges halos: 0 = nothing partitioned, 1 = ncclSend/ncclRecv,
 * 2 = peer-to-peer pushes into the neighbours' mapped ghost buffers (default; set
 * B200KS_HALO=nccl to force 1).  B200KS_FORCE_PARTITION=z|t|zt makes b200ks_create_dist treat
 * an unsplit direction as partitioned with the rank as its own neighbour (testing/profiling
 * of the halo path on fewer GPUs). */
int b200ks_halo_mode(b200ks_ctx *ctx);

void b200ks_destroy(b200ks_ctx *ctx);

/* Upload fat and long links (MILC order, su3_matrix[4*V]) and re-lay them out on the
 * device.  Replaces the implicit link refresh of the QUDA seam
 * (generic_ks/d_congrad5_fn_gpu.c:121-126).
 * long_recon selects the device storage of the long (Naik) links:
 *   18  full matrices;
 *   14  two rows + one complex factor f with row3 = f*conj(row1 x row2) -- exact for links
 *       that are (real scalar) x U(3), which HISQ/asqtad long links are
 *       (generic_ks/fermion_links_hisq_load_milc.c builds them as c_naik * W W W from the
 *       unitarised W links).  Every link is tested on load; a misfit above 1e-13 (double
 *       hosts) is an error;
 *    0  automatic: 14 when every link passes that test, else 18.
 * (Thirteen reals -- QUDA's reconstruct-13 -- would need a sincos per link and breaks the
 * 16-byte word the kernels load; 14 keeps 89 % of its traffic saving.) */
int b200ks_load_links(b200ks_ctx *ctx, const void *fat, const void *lng, int host_prec,
                      int long_recon);
/* 64-bit content fingerprint of a host array (threaded, memory-bandwidth bound; every word passes a
 * non-linear mixing step, so whole time slices of sign flips -- boundary_twist_fn -- are seen). */
unsigned long long b200ks_fingerprint(const void *host, size_t bytes);
/* Keeps the device links in step with MILC's host arrays; what both shims call before every solve
 * instead of b200ks_load_links.  Replaces the refresh logic of the QUDA seam
 * (generic_ks/d_congrad5_fn_gpu.c:121-126: fn pointer + notify flag), which misses the in-place edits of
 * boundary_twist_fn (generic_ks/fermion_links_fn_twist_milc.c:318-400).
 *   changed_hint != 0, other arrays or precision than at the last sync, or nothing loaded: upload now
 *   (fingerprints taken beside the upload); returns 1.
 *   otherwise, by mode:
 *     0  trust the hint (the reference's own behaviour); returns 0
 *     1  upload anyway; returns 1
 *     2  verify in the background: host threads fingerprint both arrays WHILE the next host-buffer call
 *        (b200ks_congrad, _congrad_block, _multicg, _dslash, _mat_invert_uml, _multicg_rational) computes on
 *        the resident links; that call joins the verification before it hands anything back and, if the
 *        arrays did change, uploads them and repeats its computation.  Returns 0.
 *     3  verify now (blocking); returns 1 if the links had changed and were uploaded. */
int b200ks_links_sync(b200ks_ctx *ctx, const void *fat, const void *lng, int host_prec, int changed_hint, int mode);
int b200ks_links_sync_stats(b200ks_ctx *ctx, long long *uploads, long long *verifications);
/* Storage chosen for the long links (7 or 9 complex per link) and the worst misfit measured
 * by the load-time test (-1 when long_recon == 18 skipped it). */
int b200ks_long_link_info(b200ks_ctx *ctx, int *ncomplex_per_link, double *misfit);

/* dest(parity sites) = D src.  Only `parity` sites of dest are written; src == dest is
 * legal for EVEN/ODD.  Replaces dslash_fn_field (generic_ks/dslash_fn.c:306-356,
 * generic_ks/dslash_fn_dblstore.c:285-304). */
int b200ks_dslash(b200ks_ctx *ctx, const void *src, void *dest, int parity, int host_prec);

/* Single-mass CG: (4 m^2 - D D) dest = src on args->parity; dest = initial guess in,
 * solution out.  Replaces ks_congrad_parity_gpu / qudaInvert
 * (generic_ks/d_congrad5_fn_gpu.c:35-172) with the CPU algorithm's semantics
 * (generic_ks/d_congrad5_fn_milc.c:60-407).  Returns iterations (>= 0) or an error. */
int b200ks_congrad(b200ks_ctx *ctx, const void *src, void *dest, double mass,
                   const b200ks_invert_args *args, b200ks_invert_result *res, int host_prec);

/* Block (multi-right-hand-side) single-mass CG: nsrc independent systems
 * (4 m^2 - D D) dest[k] = src[k] with one mass and one set of links, solved K <= 4 at a time
 * by a stencil that loads every link once for the K colour vectors (link traffic per
 * right-hand side / K).  Replaces ks_congrad_block_parity_gpu / qudaInvertMsrc
 * (generic_ks/d_congrad5_fn_gpu.c:175-312); the CPU reference is a loop of single solves
 * (generic_ks/d_congrad5_fn_milc.c:409-417).  Per right-hand side the result is that of
 * b200ks_congrad: with mixed_precision 0 the same arithmetic, iteration counts and restarts
 * (a right-hand side that stops early idles until the others have); with mixed_precision != 0
 * single-precision Krylov vectors with joint reliable updates.  res has nsrc entries
 * (device_seconds = time of the group of <= 4 the source was solved in).  Returns the total
 * number of iterations like the reference loop.  Multi-GPU contexts run it K-wide as well (one halo exchange for
 * the K inputs); the Fermilab relative residual runs the loop. */
int b200ks_congrad_block(b200ks_ctx *ctx, int nsrc, const void *const *src, void *const *dest,
                         double mass, const b200ks_invert_args *args, b200ks_invert_result *res,
                         int host_prec);

/* Multi-shift CG: (offset_j - D D) psim[j] = src for all j.  psim[j] are zeroed first.
 * Replaces ks_multicg_offset_field_gpu / qudaMultishiftInvert
 * (generic_ks/ks_multicg_offset_gpu.c:38-252) with the CPU algorithm's semantics
 * (generic_ks/ks_multicg_offset.c:63-505).  res has num_offsets entries. */
int b200ks_multicg(b200ks_ctx *ctx, const void *src, void *const *psim, const double *offsets,
                   int num_offsets, const b200ks_invert_args *args, b200ks_invert_result *res,
                   int host_prec);

/* ---- device-resident solve sequences ----------------------
 * b200ks_mat_invert_uml: dst = M^-1 src, M = D + 2m, on BOTH parities for nsrc sources, the
 *   sequence of mat_invert_uml_field / mat_invert_block_uml (generic_ks/mat_invert.c:328-402,
 *   409-475): tmp = M^+ src; even solve (M^+ M) dst_e = tmp_e starting from dst_e; odd sites
 *   reconstructed, dst_o = (src_o - D_oe dst_e)/2m; odd solve from that guess.  One upload of
 *   src and dst, one download of dst per source; the sources go through the block solver four at
 *   a time.  args->parity is ignored; res[2k], res[2k+1] = even and odd solve of source k;
 *   returns the total number of iterations (MILC: qic->final_iters = even + odd).
 * b200ks_multicg_rational: multi-shift solve + what its RHMC callers do next, on the device:
 *   fill_other != 0: psim[j](other parity) = D psim[j], both parities returned -- the input the
 *     fermion force wants (ks_imp_rhmc/update_h_rhmc.c:75-86, one dslash_field per shift);
 *   residues != NULL: dest(parity) = residues[0] src + sum_j residues[j+1] psim[j], ks_rateval
 *     (ks_imp_rhmc/ks_ratinv.c:121-138); psim may then be NULL and only dest travels back. */
int b200ks_mat_invert_uml(b200ks_ctx *ctx, int nsrc, const void *const *src, void *const *dst, double mass,
                          const b200ks_invert_args *args, b200ks_invert_result *res, int host_prec);
int b200ks_mat_invert_uml_dev(b200ks_ctx *ctx, int nsrc, const int *vsrc, const int *vdst, double mass,
                              const b200ks_invert_args *args, b200ks_invert_result *res);
int b200ks_multicg_rational(b200ks_ctx *ctx, const void *src, void *const *psim, void *dest,
                            const double *offsets, const double *residues, int num_offsets, int fill_other,
                            const b200ks_invert_args *args, b200ks_invert_result *res, int host_prec);

/* ---- low-mode deflation with eigenvectors resident in HBM ------
 * Replaces deflate() + project_out() (generic_ks/mat_invert.c:131-183), which give the CGs of
 * mat_invert_uml_field / mat_invert_cg_field (:186-257,328-402; qic->deflate) the exact solution in the
 * span of the low modes as their starting point: on the sites of one parity
 *     dst <- dst - sum_j v_j <v_j|dst> + sum_j v_j <v_j|src> / (eigval_j + 4 m^2) .
 * b200ks_eig_set: declares nvecs device vectors (b200ks_vec_create / b200ks_vec_upload, both parities
 *   filled, orthonormal on each parity: MILC's eigVec[j]) and their eigenvalues of -D_eo D_oe (MILC's
 *   eigVal[j]) as the context's low-mode set; nvecs = 0 drops it.  The vectors stay in HBM (48 B per
 *   site and parity each) and cannot be freed while in the set.  use_in_uml != 0: b200ks_mat_invert_uml
 *   and _dev deflate their trial solutions before the even and before the odd solve, as the
 *   reference does with qic->deflate set.
 * b200ks_deflate_dev: the update above for device vectors, parity EVEN or ODD.
 * The reference removes the modes from dst one after the other; the batch form used here is identical
 * for orthonormal vectors.  Single-GPU contexts. */
int b200ks_eig_set(b200ks_ctx *ctx, int nvecs, const int *vecs, const double *eigval, int use_in_uml);
int b200ks_eig_count(b200ks_ctx *ctx);
int b200ks_eig_use_in_uml(b200ks_ctx *ctx, int on);   /* qic->deflate of the next UML sequences */
int b200ks_deflate_dev(b200ks_ctx *ctx, int vsrc, int vdst, double mass, int parity);

/* ---- eigCG: solve and harvest low modes in the same Krylov space ----------
 * Replaces ks_eigCG_parity / ks_inc_eigCG_parity / calc_eigenpairs (generic_ks/inc_eigcg.c:377-850, 851-950,
 * 282-300; A. Stathopoulos and K. Orginos, arXiv:0707.0131), what mat_invert_uml_field calls for its even-site
 * solve when MILC is built with EIGMODE = EIGCG (generic_ks/mat_invert.c:361-363).
 * b200ks_eigcg_init: starts an incremental sequence with MILC's eigcg_params (include/imp_ferm_links.h:410-416):
 *   a search window of m vectors, nvecs Ritz pairs harvested per solve, at most nvecs_max accumulated.  The
 *   window and the accumulated vectors live in HBM, one parity half each (24 B x lattice volume per vector).
 * b200ks_inc_eigcg[_dev]: one solve of the sequence: the trial solution is first improved by the accumulated
 *   vectors, x += U (H + 4 m^2)^-1 U^+ (b - A x) (initCG); then the context's pure-double CG runs -- iteration
 *   for iteration the arithmetic of b200ks_congrad -- while its coefficients build the Lanczos matrix and its
 *   normalised residuals fill the window, which is compressed to the 2 nvecs Ritz vectors of T_m and T_{m-1}
 *   whenever it is full; the nvecs lowest Ritz vectors are orthogonalised against the accumulated set and
 *   H = -U^+ D^2 U is extended.  Returns iterations like b200ks_congrad; res likewise.
 * b200ks_eigcg_pairs: Rayleigh-Ritz on everything accumulated; eigval (may be NULL, room for nmax_out) receives
 *   the Ritz values of -D_eo D_oe in ascending order; returns their number.
 * b200ks_eigcg_count / _vec_download / _H: what has been accumulated, for MILC's eigVec[] / eigcgp->H.
 * Single-GPU contexts. */
int b200ks_eigcg_init(b200ks_ctx *ctx, int m, int nvecs, int nvecs_max);
int b200ks_inc_eigcg(b200ks_ctx *ctx, const void *src, void *dest, double mass, const b200ks_invert_args *args,
                     b200ks_invert_result *res, int host_prec);
int b200ks_inc_eigcg_dev(b200ks_ctx *ctx, int vsrc, int vdest, double mass, const b200ks_invert_args *args,
                         b200ks_invert_result *res);
int b200ks_eigcg_pairs(b200ks_ctx *ctx, double *eigval, int nmax_out);
int b200ks_eigcg_count(b200ks_ctx *ctx);
int b200ks_eigcg_vec_download(b200ks_ctx *ctx, int j, void *host, int host_prec);
int b200ks_eigcg_hmatrix(b200ks_ctx *ctx, double *H_out);

/* ---- meson tie-ups: two propagators -> momentum-projected time-slice correlators ----
 * Replaces the site loops of ks_meson_cont_mom (generic_ks/ks_meson_mom.c:160-437) for ONE sink spin-taste
 * assignment g of its table:
 *     corr[t][p] = sum_{x in time slice t} s(x - r0) <antiquark(x) | quark(x)> ftfact_p(x - r0),   t < nt, p < nmom
 *   <a|b> = su3_dot(a, b) (:349-363); ftfact_p = product over x, y, z of cos / i sin / exp(i .) of
 *   2 pi (x_d - r0_d) mom[3p+d] / n_d chosen by mom_parity[3p+d] = B200KS_EVEN / _ODD / _EVENANDODD (MILC's
 *   q_momstore[p], q_parity[p]; ff(), :137-157); s = the site sign of a LOCAL sink operator applied to the
 *   antiquark (local(), generic_ks/spin_taste_ops.c:172-263): spin = its gamma bits 0..15 (gamma_hex of the
 *   operator's spin index: pion5 = 15, pion05 = 0, rhox = 1, rhoy = 2, rhoz = 4, rhox0 = 9, ...), or spin = -1 when
 *   the caller has applied the sink operator to `antiquark` itself (the one-link and FN-shifted operators, which
 *   MILC builds from fat-link shifts on the host).
 * corr: nt x nmom complex numbers (re, im), OVERWRITTEN; nt is the GLOBAL time extent.  The correlator phase and
 * factor (norm_v, :100-131) and the accumulation into prop[corr_index][t] stay with the caller (csrc_milc/milc_shim.c
 * ks_meson_cont_mom_gpu).  Both propagators are read once whatever nmom is (nmom <= 128 per call).
 * _dev: device vectors (b200ks_vec_create; e.g. the solutions of b200ks_mat_invert_uml_dev, never leaving HBM).
 * Multi-GPU contexts (b200ks_create_multi): every member contracts its sub-lattice, the shares are added on the
 * host in member order. */
int b200ks_meson_mom_dev(b200ks_ctx *ctx, int vantiquark, int vquark, int spin, const int *r0, int nmom, const int *mom,
                         const char *mom_parity, double *corr);
int b200ks_meson_mom(b200ks_ctx *ctx, const void *antiquark, const void *quark, int host_prec, int spin, const int *r0,
                     int nmom, const int *mom, const char *mom_parity, double *corr);

/* ---- fermion-link construction ---------------------------
 * Links are su3_matrix[4*V] in MILC order (link[4*i+dir]) with KS phases and boundary signs
 * in (phases_in = 1); path_coeff = {one_link, naik, three_staple, five_staple, seven_staple,
 * lepage} (asqtad_coeffs_t, include/ks_action_paths.h:23-30).  Double precision on the device
 * whatever host_prec is.  Single-GPU contexts.
 *
 * b200ks_ks_links: fatlink = smeared inlink, longlink (may be NULL) = naik * U U U.
 *   Replaces load_fatlinks_cpu + load_lnglinks (generic_ks/fermion_links_fn_load_milc.c:45-275)
 *   = qudaLoadKSLink (generic_ks/fermion_links_fn_load_gpu.c:36,67).
 * b200ks_unitarized_links: vlink (may be NULL) = smeared inlink, wlink = its U(3) projection
 *   W = V (V^+ V)^-1/2.  Replaces load_V_from_U + load_Y_from_V with UNITARIZE_ANALYTIC
 *   (generic_ks/fermion_links_hisq_load_milc.c:118-340, su3_mat_op.c:828-1205)
 *   = qudaLoadUnitarizedLink (fermion_links_fn_load_gpu.c:108).  *nsvd (may be NULL) returns how
 *   many links took the SVD branch (HISQ_REUNIT_ALLOW_SVD; thresholds 1e-8 as in
 *   ks_imp_rhmc/Make_template:204-206, or B200KS_REUNIT_ALLOW_SVD / _SVD_REL_ERROR /
 *   _SVD_ABS_ERROR in the environment).
 * b200ks_hisq_links: the whole chain U -> V -> W -> (fat, long) of create_hisq_links_milc
 *   (fermion_links_hisq_load_milc.c:684-713, one Naik epsilon) with the intermediate fields
 *   resident: coeff1 = level-1 (fat7), coeff2 = level-2 coefficients; any output may be NULL. */
int b200ks_ks_links(b200ks_ctx *ctx, const double *path_coeff, const void *inlink, void *fatlink,
                    void *longlink, int host_prec);
int b200ks_unitarized_links(b200ks_ctx *ctx, const double *path_coeff, const void *inlink, void *vlink,
                            void *wlink, int host_prec, long long *nsvd);
int b200ks_hisq_links(b200ks_ctx *ctx, const double *coeff1, const double *coeff2, const void *inlink,
                      void *vlink, void *wlink, void *fatlink, void *longlink, int host_prec,
                      long long *nsvd);
/* Benchmark face: the chain on Haar-random thin links generated on the device, `reps` times,
 * CUDA-event milliseconds per chain; b200ks_hisq_links_fetch reads back field `which` of the last
 * chain (0 input, 1 V, 2 W, 3 fat, 4 long) for the CPU comparison. */
int b200ks_hisq_links_time(b200ks_ctx *ctx, const double *coeff1, const double *coeff2,
                           unsigned long long seed, int reps, double *ms_per_chain, long long *nsvd);
int b200ks_hisq_links_fetch(b200ks_ctx *ctx, int which, void *host, int host_prec);

/* ---- HISQ fermion force ---------------------------------------
 * momentum = eps * (traceless anti-Hermitian force) for S = sum_j res_j |D_oe X_j|^2, the
 * increment MILC adds to its momenta.  Replaces fn_fermion_force_multi_hisq_wrapper_mx
 * (generic_ks/fermion_force_hisq_multi.c:1183-1476) = qudaHisqForce (:2169-2290), whose argument
 * conventions it keeps:
 *   coeff[2*j], coeff[2*j+1]  one-hop (2 res_j) and three-hop (naik * 2 res_j) weights of term j < nterms
 *   num_naik_terms            several Naik epsilons (:2196-2222): the LAST num_naik_terms of the nterms fields were
 *                             solved with a Naik epsilon; term i of those (field multi_x[nterms - num_naik_terms
 *                             + i]) has its extra weights eps_k c1' 2 res and eps_k c3' 2 res (c1', c3' the
 *                             one-link + Naik table) in coeff[2*(nterms+i)], coeff[2*(nterms+i)+1]; 0 = none
 *   multi_x[j]                su3_vector[V]: solution on the even sites, D solution on the odd sites
 *   level2_coeff, fat7_coeff  the six path coefficients of the two smearing levels
 *   wlink, vlink, ulink       W (unitarised), V (fat7) and U (thin, phases in): su3_matrix[4*V]
 *   force_filter              the reference's HISQ_FORCE_FILTER (generic_ks/su3_mat_op.c:1680-1734; 5e-5 in
 *                             ks_imp_rhmc's build, QudaHisqParams_t.force_filter at the seam): a link whose
 *                             V^+V has an eigenvalue below it gets the derivative of V (V^+V + filter)^-1/2;
 *                             0 = the unregularised derivative
 *   momentum                  out: anti_hermitmat[4*V], 10 reals each (include/su3.h)
 * Computed as the reverse-mode derivative of the link construction (csrc/force.cuh); double on the
 * device whatever host_prec is.  Single-GPU contexts. */
int b200ks_hisq_force(b200ks_ctx *ctx, int nterms, int num_naik_terms, const double *coeff, const void *const *multi_x,
                      const double *level2_coeff, const double *fat7_coeff, const void *wlink,
                      const void *vlink, const void *ulink, double eps, double force_filter, void *momentum,
                      int host_prec);

/* ---- device-resident interface (benchmarks, resident solve sequences) -------------- */

/* Device colour-vector fields of one context, double precision, both parities. */
int b200ks_vec_create(b200ks_ctx *ctx);             /* returns handle >= 0 */
int b200ks_vec_free(b200ks_ctx *ctx, int vec);
int b200ks_vec_upload(b200ks_ctx *ctx, int vec, const void *host, int parity, int host_prec);
int b200ks_vec_download(b200ks_ctx *ctx, int vec, void *host, int parity, int host_prec);
int b200ks_vec_zero(b200ks_ctx *ctx, int vec, int parity);
int b200ks_vec_gaussian(b200ks_ctx *ctx, int vec, int parity, unsigned long long seed);
int b200ks_vec_norm2(b200ks_ctx *ctx, int vec, int parity, double *out);

/* Synthetic HISQ-like links generated on the device (for volumes whose host copy
 * would not fit in host RAM, SURVEY.md section 7 "Host memory at the 96^3x192 point"). */
int b200ks_links_synthetic(b200ks_ctx *ctx, unsigned long long seed, int long_recon);
/* Read the device links back in MILC host layout (local sub-lattice). */
int b200ks_links_download(b200ks_ctx *ctx, void *fat, void *lng, int host_prec);

/* prec = B200KS_PREC_DOUBLE applies the double stencil; SINGLE / HALF apply the stencil the
 * mixed solvers iterate with to device-converted copies and widen the result. */
int b200ks_dslash_dev(b200ks_ctx *ctx, int vsrc, int vdest, int parity, int prec);
int b200ks_congrad_dev(b200ks_ctx *ctx, int vsrc, int vdest, double mass,
                       const b200ks_invert_args *args, b200ks_invert_result *res);
int b200ks_multicg_dev(b200ks_ctx *ctx, int vsrc, const int *vpsim, const double *offsets,
                       int num_offsets, const b200ks_invert_args *args,
                       b200ks_invert_result *res);

int b200ks_congrad_block_dev(b200ks_ctx *ctx, int nsrc, const int *vsrc, const int *vdest, double mass,
                             const b200ks_invert_args *args, b200ks_invert_result *res);
/* D applied to nrhs (1..4) device vectors in one pass (prec DOUBLE or SINGLE), and its timing
 * probe (milliseconds per launch of the nrhs-wide stencil). */
int b200ks_dslash_block_dev(b200ks_ctx *ctx, int nrhs, const int *vsrc, const int *vdest, int parity,
                            int prec);
int b200ks_dslash_block_time(b200ks_ctx *ctx, int prec, int nrhs, int parity, int n,
                             double *ms_per_launch);

/* Times n back-to-back dslash launches (one parity) with CUDA events on the library's
 * stream; returns milliseconds per launch in *ms_per_launch. */
int b200ks_dslash_time(b200ks_ctx *ctx, int prec, int parity, int n, double *ms_per_launch);

/* Where the wall time of the last b200ks_congrad call went, seconds: out8[0] host side of the two uploads,
 * [1] solve (device time + polling), [2] what waiting for the link verification (b200ks_links_sync mode 2)
 * added after the solve, [3] download of the solution, [4] whole call, [5] passes (2 = the links had
 * changed and the solve was repeated), [6] CUDA-event time of the solve proper. */
int b200ks_call_profile(b200ks_ctx *ctx, double *out8);
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
long long b200ks_launch_count(b200ks_ctx *ctx);
/* Raw CUDA stream (cudaStream_t) the library launches on, for external event timing. */
void *b200ks_stream(b200ks_ctx *ctx);
/* Bytes of device memory held by the context. */
size_t b200ks_device_bytes(b200ks_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200KS_H */
