#define LB_FMT 6
#define LB_GETTER lb_get_fft1_small_fmt6
#include "kernels_small.inc"
