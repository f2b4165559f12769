// The network parameters as seen by the kernels.  On the device they live in __constant__ memory (one copy per
// translation unit; hpv_api.cu tracks which context's parameters each copy currently holds and refreshes it,
// stream-ordered, before a launch) so that the rolled layer loops read them through the uniform
// datapath (SASS LDCU -> uniform registers -> FFMA2 with a UR operand): no shared-memory/LSU bandwidth is spent on
// broadcasting weights to 32 lanes, which is what bounded the first version of the kernels (profiles/r01a_*).
// In the host emulation build the parameters are the plain padded array.
#pragma once
#include "hpv_types.h"

#if defined(__CUDACC__)
__constant__ __align__(16) float hpv_c_theta[HPV_CTHETA_MAX];
#endif

#if defined(__CUDA_ARCH__)
#define HPV_THETA(theta_pad) (hpv_c_theta)
#else
#define HPV_THETA(theta_pad) (theta_pad)
#endif
