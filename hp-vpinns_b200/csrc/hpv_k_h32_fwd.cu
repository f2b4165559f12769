// Kernel instantiations: padded hidden width 32, kind fwd (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h32_fwd(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<32, HPV_K_VARFWD>(k, l); }
