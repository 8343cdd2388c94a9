#define LB_FMT 0
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt0
#include "kernels_pipe.inc"
