/* oracle/ref_harness/defines.h -- TEST INFRASTRUCTURE (see lattice.h). */
#ifndef _DEFINES_H
#define _DEFINES_H
#endif
