#define LB_FMT 3
#define LB_GETTER lb_get_fft1_small_fmt3
#include "kernels_small.inc"
