#define LB_FMT 2
#define LB_LARGE_LAUNCH lb_large_launch_fmt2
#include "kernels_large.inc"
