#define LB_FMT 2
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt2
#include "kernels_pipe.inc"
