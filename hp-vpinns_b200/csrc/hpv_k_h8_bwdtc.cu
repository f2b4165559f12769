// Kernel instantiations: padded hidden width 8, kind bwdtc -- the tensor-core reverse sweep (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h8_bwdtc(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<8, HPV_K_MLPBWD_TC>(k, l); }
