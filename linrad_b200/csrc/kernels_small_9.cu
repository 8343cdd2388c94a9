#define LB_FMT 9
#define LB_GETTER lb_get_fft1_small_fmt9
#include "kernels_small.inc"
