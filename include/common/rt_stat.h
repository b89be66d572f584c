/* forwards to the single B200 rt_ant header (reference: fhe-cmplr/rtlib/include/common/rt_stat.h) */
#include "rt_ant/rt_ant.h"
