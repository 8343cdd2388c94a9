#define LB_FMT 0
#define LB_GETTER lb_get_fft1_small_fmt0
#include "kernels_small.inc"
