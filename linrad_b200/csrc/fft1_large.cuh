// fft1_large.cuh -- multi-pass (four-step) fft1 for N >= 2^15; see kernels_large.cu
#pragma once
#include "fft1_small.cuh"
