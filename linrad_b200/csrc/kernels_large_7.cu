#define LB_FMT 7
#define LB_LARGE_LAUNCH lb_large_launch_fmt7
#include "kernels_large.inc"
