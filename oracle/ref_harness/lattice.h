/* oracle/ref_harness/lattice.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Minimal application header that lets the reference's own hot-path sources
 * (generic_ks/dslash_fn_dblstore.c, d_congrad5_fn_milc.c, ks_multicg_offset.c,
 * fn_links_milc.c, generic/com_vanilla.c, layout_hyper_prime.c, make_lattice.c)
 * compile WHERE THEY LIE under /root/reference, into oracle/_ref/.  Every MILC
 * application supplies its own lattice.h (cf. ks_spectrum/lattice.h:24-124); this
 * one declares only the globals and site members those files touch.
 */
#ifndef _LATTICE_H
#define _LATTICE_H

#include "defines.h"
#include "params.h"
#include "../include/random.h"
#include "../include/io_lat.h"
#include "../include/generic_ks.h"
#include "../include/fermion_links.h"
#include "../include/su3.h"

typedef struct {
  short x, y, z, t;
  char parity;
  int index;
  int space1;
  su3_matrix link[4] ALIGNMENT;
  Real phase[4];
  anti_hermitmat mom[4] ALIGNMENT;   /* written by the fermion force (generic_ks/fermion_force_hisq_multi.c) */
} site;

#ifdef CONTROL
#define EXTERN
#else
#define EXTERN extern
#endif

EXTERN int nx, ny, nz, nt;
EXTERN int iseed;
EXTERN int niter, nrestart;
EXTERN int volume;
EXTERN params param;
EXTERN int total_iters;
EXTERN Real u0, mass;
EXTERN Real rsqmin, rsqprop;
EXTERN size_t sites_on_node;
EXTERN size_t even_sites_on_node;
EXTERN size_t odd_sites_on_node;
EXTERN int number_of_nodes;
EXTERN int this_node;
EXTERN int phases_in;
EXTERN site *lattice;

#define N_POINTERS 16
EXTERN char **gen_pt[N_POINTERS];

/* generic_ks/mat_invert.c names the application's eigenpair storage (ks_spectrum/lattice.h);
   unused here (param.eigen_param.Nvecs = 0) */
/* generic_ks/fermion_force_hisq_multi.c counts into these and reads the per-Naik-epsilon term
   counts of the RHMC application (ks_imp_rhmc/lattice.h) */
EXTERN int hisq_svd_counter;
EXTERN int hisq_force_filter_counter;
EXTERN int n_order_naik_total;
EXTERN int n_orders_naik[MAX_NAIK];
EXTERN double *eigVal;
EXTERN su3_vector **eigVec;
EXTERN su3_matrix *ape_links;   /* spin_taste_ops.c (the *ape sink operators; unused here) */

#endif /* _LATTICE_H */
