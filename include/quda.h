/* include/quda.h -- MILC's glue includes <quda.h> next to <quda_milc_interface.h>
 * (generic_ks/ks_multicg_offset_gpu.c:9-10); everything it needs is in the latter. */
#ifndef B200KS_QUDA_H
#define B200KS_QUDA_H
#include "quda_milc_interface.h"
#endif
