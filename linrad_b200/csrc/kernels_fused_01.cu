#define LB_FMT 0
#define LB_FC 1
#define LB_GETTER lb_get_fft1_fused_fmt0_fc1
#include "kernels_fused.inc"
