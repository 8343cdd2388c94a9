#define LB_FMT 2
#define LB_FC 0
#define LB_GETTER lb_get_fft1_fused_fmt2_fc0
#include "kernels_fused.inc"
