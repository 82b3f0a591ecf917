/* oracle/ref_harness/params.h -- TEST INFRASTRUCTURE (see lattice.h). */
#ifndef _PARAMS_H
#define _PARAMS_H
typedef struct { int stopflag; } params;
#endif
