// Kernel instantiations: padded hidden width 20, kind bwd (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h20_bwd(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<20, HPV_K_MLPBWD>(k, l); }

#if defined(HPV_EXP_STAMPS)
// timing experiment only (see hpv_varbwd.cuh): the stamps of the last reverse-sweep launch of this translation unit
extern "C" int hpv_exp_read_bstamps(unsigned long long* out, int n_ctas) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, hpv_exp_bstamps, sizeof(unsigned long long) * HPV_EXP_NBSTAMP * (n_ctas < 256 ? n_ctas : 256));
}
#endif
