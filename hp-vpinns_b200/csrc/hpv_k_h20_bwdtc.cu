// Kernel instantiations: padded hidden width 20, kind bwdtc -- the tensor-core reverse sweep (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h20_bwdtc(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<20, HPV_K_MLPBWD_TC>(k, l); }
