#define LB_FMT 0
#define LB_LARGE_LAUNCH lb_large_launch_fmt0
#include "kernels_large.inc"
