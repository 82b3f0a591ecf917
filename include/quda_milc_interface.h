/* include/quda_milc_interface.h -- route-2 drop-in boundary of libb200ks.
 *
 * MILC selects its GPU solver at compile time (-DUSE_CG_GPU, include/imp_ferm_links.h:73-76,
 * 238-239) and its glue objects (generic_ks/d_congrad5_fn_gpu.c, ks_multicg_offset_gpu.c,
 * dslash_fn.c:302-344, generic/milc_to_quda_utilities.c, generic/make_lattice.c:25-28)
 * include a header of THIS NAME and call the entry points declared below.  The original
 * header belongs to the external QUDA library (lattice/quda, not vendored in the reference,
 * version unpinned: Makefile:421 `QUDA_HOME ?= ${HOME}/quda`); this file is written from the
 * reference's call sites only, which fix every name, field and argument order used here.
 *
 * Building MILC with
 *     make PRECISION=2 WANTQUDA=true WANT_FN_CG_GPU=true QUDA_HOME=<dir holding include/ lib/>
 * therefore links the unmodified MILC tree against libb200ks.
 * Exactly these symbols are required (link-probed): qudaInit,
 * qudaSetMPICommHandle, qudaFinalize, qudaAllocatePinned, qudaFreePinned, qudaInvert,
 * qudaInvertMsrc, qudaMultishiftInvert, qudaDslash (+ qudaMomAction for ks_imp_rhmc);
 * WANT_FL_GPU=true additionally binds qudaLoadKSLink and qudaLoadUnitarizedLink, WANT_FF_GPU=true
 * qudaHisqParamsInit and qudaHisqForce.
 */
#ifndef QUDA_MILC_INTERFACE_H
#define QUDA_MILC_INTERFACE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* generic/milc_to_quda_utilities.c:15-25 */
typedef enum QudaVerbosity_s { QUDA_SILENT, QUDA_SUMMARIZE, QUDA_VERBOSE, QUDA_DEBUG_VERBOSE } QudaVerbosity;

/* generic_ks/d_congrad5_fn_gpu.c:95-102 */
typedef enum QudaParity_s { QUDA_EVEN_PARITY = 0, QUDA_ODD_PARITY, QUDA_INVALID_PARITY } QudaParity;

/* generic/milc_to_quda_utilities.c:30-33 */
typedef struct {
  const int *latsize;  /* nx, ny, nz, nt of the whole lattice */
  const int *machsize; /* logical machine grid */
  int device;          /* CUDA device ordinal */
} QudaLayout_t;

typedef struct {
  QudaVerbosity verbosity;
  QudaLayout_t layout;
} QudaInitArgs_t;

/* generic_ks/d_congrad5_fn_gpu.c:95-134, ks_multicg_offset_gpu.c:165-201, dslash_fn.c:322-341 */
typedef struct {
  int max_iter;             /* qic->max * qic->nrestart */
  QudaParity evenodd;
  int mixed_precision;      /* 0 none, 1 double/single (HALF_MIXED), 2 down to half (MAX_MIXED) */
  double boundary_phase[4];
  double tadpole;
  double naik_epsilon;
} QudaInvertArgs_t;

/* include/generic_quda.h:14-27 */
typedef struct {
  void *site;
  void *link;
  size_t link_offset;
  void *mom;
  size_t mom_offset;
  size_t size;
} QudaMILCSiteArg_t;

void qudaInit(QudaInitArgs_t input);
void qudaSetMPICommHandle(void *mycomm);
void qudaFinalize(void);

/* generic/make_lattice.c:25-28,68 ; include/generic_quda.h:45,79 */
void *qudaAllocatePinned(size_t bytes);
void qudaFreePinned(void *ptr);
void *qudaAllocateManaged(size_t bytes);
void qudaFreeManaged(void *ptr);

/* generic_ks/d_congrad5_fn_gpu.c:136-148.  *num_iters == -1 on entry: links changed, refresh
 * the device copy (:121-126). */
void qudaInvert(int external_precision, int quda_precision, double mass, QudaInvertArgs_t inv_args,
                double target_residual, double target_fermilab_residual, const void *const milc_fatlink,
                const void *const milc_longlink, void *source, void *solution, double *const final_residual,
                double *const final_fermilab_residual, int *num_iters);

/* generic_ks/d_congrad5_fn_gpu.c:268-281 */
void qudaInvertMsrc(int external_precision, int quda_precision, double mass, QudaInvertArgs_t inv_args,
                    double target_residual, double target_fermilab_residual, const void *const fatlink,
                    const void *const longlink, void **sourceArray, void **solutionArray,
                    double *const final_residual, double *const final_fermilab_residual, int *num_iters,
                    int num_src);

/* generic_ks/ks_multicg_offset_gpu.c:203-217 */
void qudaMultishiftInvert(int external_precision, int precision, int num_offsets, double *const offset,
                          QudaInvertArgs_t inv_args, const double *target_residual,
                          const double *target_fermilab_residual, const void *const milc_fatlink,
                          const void *const milc_longlink, void *source, void **solutionArray,
                          double *const final_residual, double *const final_fermilab_residual, int *num_iters);

/* generic_ks/dslash_fn.c:336-341 */
void qudaDslash(int external_precision, int quda_precision, QudaInvertArgs_t inv_args,
                const void *const milc_fatlink, const void *const milc_longlink, void *source, void *solution,
                int *num_iters);

/* ks_imp_rhmc/d_action_rhmc.c:102-105: sum over sites and directions of |mom|^2 - 4 */
double qudaMomAction(int precision, QudaMILCSiteArg_t *arg);

/* ---- fermion-link construction (-DUSE_FL_GPU: make WANT_FL_GPU=true, Makefile:453-455) ------
 * generic_ks/fermion_links_fn_load_gpu.c:18-123.  path_coeff = {one_link, naik, three_staple,
 * five_staple, seven_staple, lepage}; links are su3_matrix[4*sites_on_node] in MILC order with
 * KS phases and boundary signs in. */
typedef struct {
  int su3_source;  /* the incoming links are SU(3) (only a hint; not used here) */
} QudaFatLinkArgs_t;

/* fatlink = smeared inlink (load_fatlinks_cpu), longlink (may be NULL) = naik * three-link product
 * (load_lnglinks); called by load_fatlinks_gpu / load_fatlonglinks_gpu. */
void qudaLoadKSLink(int precision, QudaFatLinkArgs_t fatlink_args, const double path_coeff[6], void *inlink,
                    void *fatlink, void *longlink);
/* fatlink (may be NULL) = smeared inlink, ulink = its U(3) projection (u3_unitarize_analytic);
 * called by load_hisq_aux_links_gpu with the level-1 (fat7) coefficients. */
void qudaLoadUnitarizedLink(int precision, QudaFatLinkArgs_t fatlink_args, const double path_coeff[6], void *inlink,
                            void *fatlink, void *ulink);

/* ---- HISQ fermion force (-DUSE_FF_GPU: make WANT_FF_GPU=true, Makefile:458-461) ------------
 * generic_ks/fermion_force_hisq_multi.c:2169-2290. */
typedef struct {
  int reunit_allow_svd;
  int reunit_svd_only;
  double reunit_svd_abs_error;
  double reunit_svd_rel_error;
  double force_filter;
} QudaHisqParams_t;

void qudaHisqParamsInit(QudaHisqParams_t hisq_params);
/* momentum (anti_hermitmat[4*sites], precision reals) = dt * force of num_terms terms; coeff[t][0]
 * one-hop and coeff[t][1] three-hop weights; quark_field[t] su3_vector[sites] with both parities
 * filled.  num_naik_terms > 0 (several Naik epsilons): the last num_naik_terms fields carry extra
 * weights coeff[num_terms + i] (fermion_force_hisq_multi.c:2196-2222). */
void qudaHisqForce(int precision, int num_terms, int num_naik_terms, double dt, double **coeff, void **quark_field,
                   const double level2_coeff[6], const double fat7_coeff[6], const void *const w_link,
                   const void *const v_link, const void *const u_link, void *const milc_momentum);

#ifdef __cplusplus
}
#endif
#endif /* QUDA_MILC_INTERFACE_H */
