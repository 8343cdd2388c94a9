#define LB_FMT 3
#define LB_LARGE_LAUNCH lb_large_launch_fmt3
#include "kernels_large.inc"
