#define LB_FMT 2
#define LB_GETTER lb_get_fft1_small_fmt2
#include "kernels_small.inc"
