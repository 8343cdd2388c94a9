#define LB_FMT 1
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt1
#include "kernels_pipe.inc"
