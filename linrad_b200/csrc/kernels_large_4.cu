#define LB_FMT 4
#define LB_LARGE_LAUNCH lb_large_launch_fmt4
#include "kernels_large.inc"
