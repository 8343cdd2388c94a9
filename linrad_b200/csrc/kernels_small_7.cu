#define LB_FMT 7
#define LB_GETTER lb_get_fft1_small_fmt7
#include "kernels_small.inc"
