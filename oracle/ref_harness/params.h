/* oracle/ref_harness/params.h -- TEST INFRASTRUCTURE (see lattice.h). */
#ifndef _PARAMS_H
#define _PARAMS_H
#include "../include/generic_quark_types.h"
#include "../include/imp_ferm_links.h"
/* the members the compiled reference sources touch: generic_ks/mat_invert.c reads
   param.eigen_param.Nvecs (0 here: no deflation) */
typedef struct {
  int stopflag;
  ks_eigen_param eigen_param;
} params;
#endif
