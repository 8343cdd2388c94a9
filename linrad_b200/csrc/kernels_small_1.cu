#define LB_FMT 1
#define LB_GETTER lb_get_fft1_small_fmt1
#include "kernels_small.inc"
