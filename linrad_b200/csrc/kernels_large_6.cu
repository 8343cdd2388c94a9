#define LB_FMT 6
#define LB_LARGE_LAUNCH lb_large_launch_fmt6
#include "kernels_large.inc"
