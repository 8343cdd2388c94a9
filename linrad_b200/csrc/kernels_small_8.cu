#define LB_FMT 8
#define LB_GETTER lb_get_fft1_small_fmt8
#include "kernels_small.inc"
