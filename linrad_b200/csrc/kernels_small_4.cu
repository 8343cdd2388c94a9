#define LB_FMT 4
#define LB_GETTER lb_get_fft1_small_fmt4
#include "kernels_small.inc"
