#define LB_FMT 2
#define LB_FC 2
#define LB_GETTER lb_get_fft1_fused_fmt2_fc2
#include "kernels_fused.inc"
