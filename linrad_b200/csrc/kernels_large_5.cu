#define LB_FMT 5
#define LB_LARGE_LAUNCH lb_large_launch_fmt5
#include "kernels_large.inc"
