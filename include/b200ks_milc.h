/* include/b200ks_milc.h -- route-1 drop-in boundary: MILC's own solver symbols.
 *
 * Declares, with MILC's names and argument lists (include/imp_ferm_links.h:73-93,238-246,
 * 288-294), the functions MILC's generic_ks callers link against when built with
 * -DUSE_CG_GPU:
 *     ks_congrad_parity_gpu, ks_congrad_block_parity_gpu   (generic_ks/d_congrad5_fn_gpu.c)
 *     ks_multicg_offset_field_gpu, get_fn_last, set_fn_last (generic_ks/ks_multicg_offset_gpu.c)
 *     dslash_fn_field                                        (generic_ks/dslash_fn.c:306-344)
 * libb200ks_milc.so implements them directly on the b200ks C ABI (no quda* layer in
 * between).  Inside a MILC tree the same source (milc_qcd_b200/csrc_milc/milc_shim.c) is
 * compiled against MILC's own headers instead of the mirror types below (-DB200KS_IN_MILC),
 * replacing d_congrad5_fn_gpu.o and ks_multicg_offset_gpu.o in Make_template_combos:243-297.
 *
 * The mirror types reproduce the reference's layouts for a standalone build:
 *   su3_vector, su3_matrix        include/milc_datatypes.h:48-56 (via include/su3.h)
 *   quark_invert_control, ks_param include/generic_quark_types.h:131-139,167-190
 *   fn_links_t                    include/fn_links.h:12-20
 * MILC_PRECISION (1|2) selects Real exactly as include/precision.h:5-13 does.
 */
#ifndef B200KS_MILC_H
#define B200KS_MILC_H

#ifndef B200KS_IN_MILC

#ifndef MILC_PRECISION
#define MILC_PRECISION 2
#endif
#if MILC_PRECISION == 1
typedef float Real;
#else
typedef double Real;
#endif

typedef struct { Real real; Real imag; } b200ks_complex;
#define B200KS_MILC_COMPLEX b200ks_complex   /* MILC's `complex` (include/complex.h:177-180) */
typedef struct { b200ks_complex c[3]; } su3_vector;
typedef struct { b200ks_complex e[3][3]; } su3_matrix;

#define EVEN 0x02
#define ODD 0x01
#define EVENANDODD 0x03
#define MAXFILENAME 256

typedef struct {
  Real mass;
  Real charge;
  Real offset;
  Real residue;
  int naik_term_epsilon_index;
  int charge_index;
  Real naik_term_epsilon;
} ks_param;

enum inv_type { MGTYPE, CGTYPE };

typedef struct {
  int prec;
  int min;
  int max;
  int nrestart;
  int parity;
  int start_flag;
  int nsrc;
  int deflate;
  Real resid;
  Real relresid;
  Real mixed_rsq;
  Real final_rsq;
  Real final_relrsq;
  Real size_r;
  Real size_relr;
  int converged;
  int final_iters;
  int final_restart;
  enum inv_type inv_type;
  char mgparamfile[MAXFILENAME];
} quark_invert_control;

typedef struct {
  void *phase; /* link_phase_info_t * */
  su3_matrix *fat;
  su3_matrix *lng;
  su3_matrix *fatback;
  su3_matrix *lngback;
  double eps_naik;
  int notify_quda_new_links;
} fn_links_t;
typedef fn_links_t imp_ferm_links_t;

/* include/imp_ferm_links.h:410-416 */
typedef struct { double real; double imag; } b200ks_double_complex;
typedef struct {
  int m;          /* Number of vectors kept for the Lanczos part before restart */
  int Nvecs;      /* Number of eigenpairs computed per inversion */
  int Nvecs_curr; /* Number of eigenpairs currently computed */
  int Nvecs_max;  /* Maximum number of eigenpairs computed in entire incremental eigCG */
  b200ks_double_complex *H; /* H = -U^+ Dslash^2 U, column-major with leading dimension Nvecs_max */
} eigcg_params;

/* the globals a MILC application owns (ks_spectrum/lattice.h:63-124); the standalone
 * library keeps its own copies, set by b200ks_milc_setup */
#ifdef __cplusplus
extern "C" {
#endif
void b200ks_milc_setup(int nx, int ny, int nz, int nt, int mixed_precision);
void b200ks_milc_finalize(void);
int b200ks_milc_total_iters(void);
struct b200ks_ctx *b200ks_milc_context(void); /* the library context behind the symbols (diagnostics) */
#ifdef __cplusplus
}
#endif

#else
#define B200KS_MILC_COMPLEX complex
#endif /* !B200KS_IN_MILC */

#ifdef __cplusplus
extern "C" {
#endif

int ks_congrad_parity_gpu(su3_vector *t_src, su3_vector *t_dest, quark_invert_control *qic, Real mass,
                          imp_ferm_links_t *fn);
int ks_congrad_block_parity_gpu(int nsrc, su3_vector **t_src, su3_vector **t_dest, quark_invert_control *qic,
                                Real mass, imp_ferm_links_t *fn);
int ks_multicg_offset_field_gpu(su3_vector *src, su3_vector **psim, ks_param *ksp, int num_offsets,
                                quark_invert_control *qic, imp_ferm_links_t *fn);
void dslash_fn_field(su3_vector *src, su3_vector *dest, int parity, fn_links_t *fn);
/* Optional (no seam in the reference: a maintainer maps the names under USE_CG_GPU, see
 * INTEGRATION.md): the UML propagator solve as one device-resident sequence, prototypes of
 * mat_invert_uml_field / mat_invert_block_uml (generic_ks/mat_invert.c:328-402,409-475). */
int mat_invert_uml_field_gpu(su3_vector *src, su3_vector *dst, quark_invert_control *qic, Real mass,
                             imp_ferm_links_t *fn);
int mat_invert_block_uml_gpu(su3_vector **src, su3_vector **dst, Real mass, int nsrc, quark_invert_control *qic,
                             imp_ferm_links_t *fn);
/* Low modes for qic->deflate in the two sequences above (generic_ks/mat_invert.c:131-183): MILC's eigVec, eigVal and
 * param.eigen_param.Nvecs, handed over once after they were read or computed; kept in HBM.  nvecs = 0 drops them. */
void b200ks_milc_set_eigenvectors(int nvecs, su3_vector **eigvec, double *eigval);
/* Incremental eigCG on the device (generic_ks/inc_eigcg.c:851-950, 282-300), MILC's prototypes
 * (include/imp_ferm_links.h:417-424).  The search window and the accumulated vectors stay in HBM; after every solve
 * the NEW vectors are copied into eigVec[] and eigcgp (Nvecs_curr, Nvecs, H) is brought up to date, so MILC code that
 * reads them afterwards (calc_eigenpairs, the eigenvector files) keeps working.  Mapped by a maintainer under
 * USE_CG_GPU like the UML sequences. */
int ks_inc_eigCG_parity_gpu(su3_vector *src, su3_vector *dest, double *eigVal, su3_vector **eigVec, eigcg_params *eigcgp,
                            quark_invert_control *qic, Real mass, imp_ferm_links_t *fn);
void calc_eigenpairs_gpu(double *eigVal, su3_vector **eigVec, eigcg_params *eigcgp, int parity);
/* Meson tie-ups on the device, prototype of ks_meson_cont_mom (generic_ks/ks_meson_mom.c:160-178; mapped by a
 * maintainer like the sequences above).  The two propagators go up once per sink spin-taste assignment and are
 * contracted there for all of its momenta; LOCAL sink operators (site signs, generic_ks/spin_taste_ops.c:172-263) are
 * applied inside the kernel, every other operator by MILC's own spin_taste_op_fn on the host first (inside a MILC
 * tree; a standalone build refuses them).  norm_v and the accumulation into prop[m][t] as in the reference. */
void ks_meson_cont_mom_gpu(B200KS_MILC_COMPLEX **prop, su3_vector *src1, su3_vector *src2, int no_q_momenta, int **q_momstore,
                           char **q_parity, int no_spin_taste_corr, int num_corr_mom[], int **corr_table, int p_index[],
                           imp_ferm_links_t *fn_src1, imp_ferm_links_t *fn_src2, int spin_taste_snk[], int meson_phase[],
                           Real meson_factor[], int corr_index[], int r0[]);
imp_ferm_links_t *get_fn_last(void);
void set_fn_last(imp_ferm_links_t *fn_last_new);

#ifdef __cplusplus
}
#endif
#endif /* B200KS_MILC_H */
