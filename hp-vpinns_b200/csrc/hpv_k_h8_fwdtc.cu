// Kernel instantiations: padded hidden width 8, kind fwdtc -- the tensor-core forward kernel (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h8_fwdtc(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<8, HPV_K_VARFWD_TC>(k, l); }
