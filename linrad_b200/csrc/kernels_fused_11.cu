#define LB_FMT 1
#define LB_FC 1
#define LB_GETTER lb_get_fft1_fused_fmt1_fc1
#include "kernels_fused.inc"
