// Kernel instantiations: padded hidden width 20, kind fwdtc -- the tensor-core forward kernel (see hpv_kernels.cuh).
#include "hpv_kernels.cuh"
cudaError_t hpv_dispatch_h20_fwdtc(const HpvKernelKey& k, const HpvLaunch& l) { return hpv_dispatch_hp<20, HPV_K_VARFWD_TC>(k, l); }

#if defined(HPV_EXP_STAMPS)
// timing experiment only (see hpv_varfwd_tc.cuh): the stamps of the last forward launch of this translation unit
extern "C" int hpv_exp_read_stamps(unsigned long long* out, int n_ctas) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, hpv_exp_stamps, sizeof(unsigned long long) * HPV_EXP_NSTAMP * (n_ctas < 1024 ? n_ctas : 1024));
}
#endif
