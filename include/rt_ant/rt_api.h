/* forwards to the single B200 rt_ant header (reference: fhe-cmplr/rtlib/include/rt_ant/rt_api.h) */
#include "rt_ant/rt_ant.h"
