#define LB_FMT 4
#define LB_PIPE_LAUNCH lb_pipe_launch_fmt4
#include "kernels_pipe.inc"
