#define LB_FMT 6
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt6
#include "kernels_pipe.inc"
