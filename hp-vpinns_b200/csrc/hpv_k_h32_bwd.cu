// Kernel instantiations: padded hidden width 32, kind bwd (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h32_bwd(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<32, HPV_K_MLPBWD>(k, l); }
