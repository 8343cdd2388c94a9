#define LB_FMT 5
#define LB_GETTER lb_get_fft1_small_fmt5
#include "kernels_small.inc"
