#define LB_FMT 5
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt5
#include "kernels_pipe.inc"
