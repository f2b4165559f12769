// __global__ wrappers around the kernel bodies and the per-width dispatch function.  Included by
// hpv_k_h{8,20,32}_{fwd,bwd,pts}.cu: one translation unit per padded hidden width and kernel kind, so
// that the build parallelises (each unit instantiates 7 derivative modes x 2 activations).
#pragma once
#include <stdint.h>
#include "hpv_launch.h"

template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__(HPV_THREADS, (HpvMode<DIM, MX, MY>::NCH >= 4 ? 1 : 2)) hpv_varfwd_kernel(const __grid_constant__ HpvVarArgs a) {
    extern __shared__ __align__(16) unsigned char hpv_smem[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem; c.emu = nullptr;
    hpv_varfwd_body<DIM, MX, MY, HP, ACT>(c, a);
}

// Tensor-core form of the forward kernel (hpv_varfwd_tc.cuh): two CTAs per SM when a CTA's accumulators and A operands
// fit 256 TMEM columns, one otherwise.
template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__(HPV_THREADS, (hpv_tc_tmem_need(HpvMode<DIM, MX, MY>::NCH, HP) <= 256 ? 2 : 1))
hpv_varfwd_tc_kernel(const __grid_constant__ HpvVarArgs a) {
    extern __shared__ __align__(128) unsigned char hpv_smem_tc[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem_tc; c.emu = nullptr;
    if constexpr (hpv_tc_supported(HpvMode<DIM, MX, MY>::NCH, HP)) hpv_varfwd_tc_body<DIM, MX, MY, HP, ACT>(c, a);
}

// Tensor-core form of the reverse sweep (hpv_varbwd_tc.cuh).
template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__(HPV_THREADS, (hpv_tc_tmem_need(HpvMode<DIM, MX, MY>::NCH, HP) <= 256 ? 2 : 1))
hpv_mlpbwd_tc_kernel(const __grid_constant__ HpvBwdArgs a) {
    extern __shared__ __align__(128) unsigned char hpv_smem_tcb[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem_tcb; c.emu = nullptr;
    if constexpr (hpv_tc_supported(HpvMode<DIM, MX, MY>::NCH, HP)) hpv_mlpbwd_tc_body<DIM, MX, MY, HP, ACT>(c, a);
}

// ... and with the weight gradients on the tensor cores as well (hpv_varbwd_tcw.cuh).
template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__(HPV_THREADS, (hpv_tcw_tmem_need(HpvMode<DIM, MX, MY>::NCH, HP, 3) <= 256 ? 2 : 1))
hpv_mlpbwd_tcw_kernel(const __grid_constant__ HpvBwdArgs a) {
    extern __shared__ __align__(128) unsigned char hpv_smem_tcw[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem_tcw; c.emu = nullptr;
    if constexpr (HP <= 24) hpv_mlpbwd_tcw_body<DIM, MX, MY, HP, ACT>(c, a);
}

// Resident CTAs per SM of a tensor-core kernel from its resources.  (cudaOccupancyMaxActiveBlocksPerMultiprocessor
// answered 1 for the 89 KB / 128-register headline instance of the forward kernel although the hardware co-schedules
// two -- ncu: block limit registers 2, shared memory 2 -- so the bound is computed here; nothing depends on
// co-residency for correctness, the figure only sizes the persistent grid.)
template <typename K>
static cudaError_t hpv_tc_resident(K k, int block, size_t smem, int tmem_cols, long long* out) {
    int n = 0;
    cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, block, smem);
    if (err != cudaSuccess) return err;
    cudaFuncAttributes fa;
    if ((err = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return err;
    int dev = 0, smem_sm = 0, regs_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    const int regs_cta = ((fa.numRegs * 32 + 255) / 256) * 256 * (block / 32);
    const int by_regs = regs_cta > 0 ? regs_sm / regs_cta : 1;
    const int by_smem = (int)((size_t)smem_sm / (smem + fa.sharedSizeBytes + 1024));
    const int own = by_regs < by_smem ? by_regs : by_smem;
    if (n < own) n = own;
    const int by_tmem = 512 / tmem_cols;
    *out = n < by_tmem ? n : by_tmem;
    return cudaSuccess;
}

// Launch bounds of the reverse sweep.  The host picks the number of warps per SM and their grouping into CTAs per
// launch (plan_bwd): the warps share nothing but the constant parameters, so a CTA is only a resource container.
// The kernel is bound by shared-memory / constant-load latency, so resident warps matter more than registers:
// with one or two channels (incl. the directional mode of the headline form) it is compiled for 512 threads = 128
// registers -- ptxas fits it with the same ~90 bytes of spills it has at 168, and 15-16 warps per SM instead of
// 11-12 took C3 from 146 to 135 us (profiles/r02o) --, with three channels for 384 threads = 168 registers (no
// spills), with four or five for 256 threads (255 registers; 168 would spill 2 KB per thread).
#ifndef HPV_BWD_THREADS_2CH
#define HPV_BWD_THREADS_2CH 512
#endif
#ifndef HPV_BWD_THREADS_3CH
#define HPV_BWD_THREADS_3CH 384
#endif
template <int DIM, int MX, int MY>
struct HpvBwdBounds {
    static constexpr int NCH = HpvMode<DIM, MX, MY>::NCH;
    static constexpr int THREADS = NCH <= 2 ? HPV_BWD_THREADS_2CH : (NCH == 3 ? HPV_BWD_THREADS_3CH : HPV_THREADS);
    static constexpr int MIN_CTAS = 1;
};

template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__((HpvBwdBounds<DIM, MX, MY>::THREADS), (HpvBwdBounds<DIM, MX, MY>::MIN_CTAS)) hpv_mlpbwd_kernel(const __grid_constant__ HpvBwdArgs a) {
    extern __shared__ __align__(16) unsigned char hpv_smem[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem; c.emu = nullptr;
    hpv_mlpbwd_body<DIM, MX, MY, HP, ACT>(c, a);
}

template <int DIM, int MX, int MY, int HP, int ACT>
__global__ void __launch_bounds__(HPV_THREADS, 1) hpv_points_kernel(const __grid_constant__ HpvPointArgs a, float* gbar_out) {
    extern __shared__ __align__(16) unsigned char hpv_smem[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem; c.emu = nullptr;
    hpv_points_body<DIM, MX, MY, HP, ACT>(c, a, gbar_out);
}

// Opt in to > 48 KB of dynamic shared memory once per kernel instantiation and device (the attribute call
// costs tens of microseconds on the host, far more than a launch).  `prepared` is the per-instantiation cache
// owned by the caller (a function-local static of hpv_do, which is templated on the full kernel identity).
template <typename K>
static cudaError_t hpv_prepare(K kernel, size_t smem, size_t* prepared) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (smem <= prepared[dev] && prepared[dev] != 0) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) prepared[dev] = smem;
    return e;
}

template <int DIM, int MX, int MY, int HP, int ACT, int KIND>
static cudaError_t hpv_do(const HpvLaunch& l) {
    cudaError_t err = cudaSuccess;
    static size_t prepared[16] = {0};
    if constexpr (KIND == HPV_K_VARFWD) {
        auto k = hpv_varfwd_kernel<DIM, MX, MY, HP, ACT>;
        if ((err = hpv_prepare(k, l.smem, prepared)) != cudaSuccess) return err;
        if (l.op == 1) {
            int n = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, l.block, l.smem);
            *l.out = n;
            return err;
        }
        if (l.op == 3) { void* p = nullptr; err = cudaGetSymbolAddress(&p, hpv_c_theta); *l.out = (long long)(uintptr_t)p; return err; }
        k<<<l.grid, l.block, l.smem, l.stream>>>(*l.var);
    } else if constexpr (KIND == HPV_K_VARFWD_TC) {
        if constexpr (!hpv_tc_supported(HpvMode<DIM, MX, MY>::NCH, HP)) {
            return cudaErrorInvalidValue;
        } else {
            auto k = hpv_varfwd_tc_kernel<DIM, MX, MY, HP, ACT>;
            if ((err = hpv_prepare(k, l.smem, prepared)) != cudaSuccess) return err;
            if (l.op == 1) {
                return hpv_tc_resident(k, l.block, l.smem, hpv_tc_tmem_cols(HpvMode<DIM, MX, MY>::NCH, HP), l.out);
            }
            k<<<l.grid, l.block, l.smem, l.stream>>>(*l.var);
        }
    } else if constexpr (KIND == HPV_K_MLPBWD_TC) {
        typedef HpvMode<DIM, MX, MY> M;
        if constexpr (!hpv_tc_supported(M::NCH, HP)) {
            return cudaErrorInvalidValue;
        } else {
            const int nhid = l.bwd->v.nhid;
            bool wg = false;
            if constexpr (HP <= 24) wg = l.wg != 0 && hpv_tcw_supported(M::NCH, HP, nhid);
            if (l.op == 5) { *l.out = wg ? 1 : 0; return cudaSuccess; }       // which variant would run
            if constexpr (HP <= 24) {
                if (wg) {
                    auto k = hpv_mlpbwd_tcw_kernel<DIM, MX, MY, HP, ACT>;
                    if (l.op == 2) { *l.out = (long long)hpv_bwd_tcw_smem(DIM, HP, M::NCH, nhid).total * 4; return cudaSuccess; }
                    static size_t prepared_w[16] = {0};
                    if ((err = hpv_prepare(k, l.smem, prepared_w)) != cudaSuccess) return err;
                    if (l.op == 1) return hpv_tc_resident(k, l.block, l.smem, hpv_tcw_tmem_need(M::NCH, HP, nhid) <= 256 ? 256 : 512, l.out);
                    return hpv_launch_pdl(k, l.grid, l.block, l.smem, l.stream, *l.bwd);
                }
            }
            auto k = hpv_mlpbwd_tc_kernel<DIM, MX, MY, HP, ACT>;
            if (l.op == 2) {
                const HpvBwdTcSmem L = hpv_bwd_tc_smem(DIM, HP, M::NCH, 1 + (M::DX ? 1 : 0) + (M::DY ? 1 : 0), nhid);
                *l.out = (long long)L.total * 4;
                return cudaSuccess;
            }
            if ((err = hpv_prepare(k, l.smem, prepared)) != cudaSuccess) return err;
            if (l.op == 1) return hpv_tc_resident(k, l.block, l.smem, hpv_tc_tmem_cols(M::NCH, HP), l.out);
            return hpv_launch_pdl(k, l.grid, l.block, l.smem, l.stream, *l.bwd);
        }
    } else if constexpr (KIND == HPV_K_MLPBWD) {
        auto k = hpv_mlpbwd_kernel<DIM, MX, MY, HP, ACT>;
        if (l.op == 2) {
            HpvBwdSmem<DIM, MX, MY, HP> L(l.bwd->v.nhid, l.block);
            *l.out = (long long)L.total * 4;
            return cudaSuccess;
        }
        if ((err = hpv_prepare(k, l.smem, prepared)) != cudaSuccess) return err;
        if (l.op == 1) {
            int n = 0;
            cudaFuncAttributes fa;
            if ((err = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return err;
            if (l.block > fa.maxThreadsPerBlock) { *l.out = 0; return cudaSuccess; }     // beyond the launch bounds
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, l.block, l.smem);
            *l.out = n;
            return err;
        }
        if (l.op == 4) {                 // threads per block the kernel was compiled for
            cudaFuncAttributes fa;
            if ((err = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return err;
            *l.out = fa.maxThreadsPerBlock;
            return cudaSuccess;
        }
        if (l.op == 3) { void* p = nullptr; err = cudaGetSymbolAddress(&p, hpv_c_theta); *l.out = (long long)(uintptr_t)p; return err; }
        return hpv_launch_pdl(k, l.grid, l.block, l.smem, l.stream, *l.bwd);
    } else {
        auto k = hpv_points_kernel<DIM, MX, MY, HP, ACT>;
        if ((err = hpv_prepare(k, l.smem, prepared)) != cudaSuccess) return err;
        if (l.op == 1) {
            int n = 0;
            err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, l.block, l.smem);
            *l.out = n;
            return err;
        }
        if (l.op == 3) { void* p = nullptr; err = cudaGetSymbolAddress(&p, hpv_c_theta); *l.out = (long long)(uintptr_t)p; return err; }
        k<<<l.grid, l.block, l.smem, l.stream>>>(*l.pts, l.gbar_out);
    }
    return cudaGetLastError();
}

template <int DIM, int MX, int MY, int HP, int KIND>
static cudaError_t hpv_do_act(const HpvKernelKey& k, const HpvLaunch& l) {
    if (k.act == HPV_ACT_TANH) return hpv_do<DIM, MX, MY, HP, HPV_ACT_TANH, KIND>(l);
    return hpv_do<DIM, MX, MY, HP, HPV_ACT_SIN, KIND>(l);
}

template <int HP, int KIND>
static cudaError_t hpv_dispatch_hp(const HpvKernelKey& k, const HpvLaunch& l) {
    if (k.dim == 1) {
        if (k.mx == 0) return hpv_do_act<1, 0, 0, HP, KIND>(k, l);
        if (k.mx == 1) return hpv_do_act<1, 1, 0, HP, KIND>(k, l);
        return hpv_do_act<1, 2, 0, HP, KIND>(k, l);
    }
    if (k.mx == 0 && k.my == 0) return hpv_do_act<2, 0, 0, HP, KIND>(k, l);
    if constexpr (KIND == HPV_K_MLPBWD || KIND == HPV_K_MLPBWD_TC) {
        if (k.dir && k.mx <= 1 && k.my <= 1) return hpv_do_act<2, 1, 0, HP, KIND>(k, l);
    }
    if (k.mx <= 1 && k.my <= 1) return hpv_do_act<2, 1, 1, HP, KIND>(k, l);
    if (k.mx == 2 && k.my <= 1) return hpv_do_act<2, 2, 1, HP, KIND>(k, l);
    return hpv_do_act<2, 2, 2, HP, KIND>(k, l);
}
