#define LB_FMT 1
#define LB_LARGE_LAUNCH lb_large_launch_fmt1
#include "kernels_large.inc"
