// Mode-independent kernels: adjoint projection (K1), partial-gradient reduction (K3), the TF1-semantics Adam
// update (tf.train.AdamOptimizer, P2D:131-132 / P1D:102-104 / ADI:191-193) and the FP32-FFMA peak probe that
// bench.py uses as the roofline denominator of the FFMA kernels.
#include <string.h>
#include "hpv_launch.h"

__global__ void __launch_bounds__(HPV_THREADS, 1) hpv_adjproj_kernel(const __grid_constant__ HpvAdjArgs a) {
    extern __shared__ __align__(16) unsigned char hpv_smem[];
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = gridDim.x;
    c.smem = hpv_smem; c.emu = nullptr;
    hpv_adjproj_body(c, a);
}

cudaError_t hpv_launch_adjproj(const HpvAdjArgs& a, int grid, size_t smem, cudaStream_t s) {
    static size_t prepared[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (smem > prepared[dev] || prepared[dev] == 0) {
        cudaError_t err = cudaFuncSetAttribute(hpv_adjproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        prepared[dev] = smem;
    }
    hpv_adjproj_kernel<<<grid, HPV_THREADS, smem, s>>>(a);
    return cudaGetLastError();
}

// CTAs 0 .. nred-1 reduce 32 gradient entries each; with has_loss the last CTA assembles the loss values.
__global__ void __launch_bounds__(256) hpv_gradreduce_kernel(const HpvGradReduceArgs a, const HpvLossArgs la, int nred) {
    __shared__ __align__(16) unsigned char smem[8 * 32 * 4];
    if ((int)blockIdx.x >= nred) {
        if (threadIdx.x < 32) hpv_losses_warp(la, threadIdx.x);
        return;
    }
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = nred;
    c.smem = smem; c.emu = nullptr;
    hpv_gradreduce_body(c, a);
}

cudaError_t hpv_launch_gradreduce(const HpvGradReduceArgs& a, const HpvLossArgs* la, cudaStream_t s) {
    const int nred = (a.n + 31) / 32;
    HpvLossArgs l0;
    memset(&l0, 0, sizeof(l0));
    hpv_gradreduce_kernel<<<nred + (la ? 1 : 0), 256, 0, s>>>(a, la ? *la : l0, nred);
    return cudaGetLastError();
}

// One CTA; thread-strided over the parameters (reference order, eps last).  Always un-pads the gradient; with
// update != 0 applies
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  lr_t = lr sqrt(1-b2^t)/(1-b1^t);  theta -= lr_t m/(sqrt(v)+eps_hat)
// (eps_hat outside the bias correction, as in TF1) in float64 master copies, refreshes the fp32 padded
// parameters the kernels read -- in global memory and in the constant-memory mirrors this context owns (the
// constant caches are invalidated at the next launch, so the following kernels see the update) -- and advances
// the step counter (single CTA: every thread has read t before thread 0 stores t + 1).
__global__ void __launch_bounds__(1024) hpv_adam_kernel(const HpvAdamArgs a) {
    const HpvAdamArgs& d = a;
    const int t = a.update ? a.step[0] + 1 : 0;
    const double b1 = a.b1, b2 = a.b2;
    const double lr_t = a.update ? (double)a.lr * sqrt(1.0 - pow(b2, (double)t)) / (1.0 - pow(b1, (double)t)) : 0.0;
    __syncthreads();
    for (int i = threadIdx.x; i <= a.n_theta; i += blockDim.x) {
        const bool is_eps = (i == a.n_theta);
        const double g = (double)(is_eps ? a.grad_pad[a.theta_pad_n] : a.grad_pad[a.pad_index[i]]);
        if (d.grad_out) d.grad_out[i] = g;
        if (!a.update) continue;
        if (is_eps && !a.train_eps) continue;
        const double m = b1 * d.m[i] + (1.0 - b1) * g;
        const double v = b2 * d.v[i] + (1.0 - b2) * g * g;
        const double th = d.theta[i] - lr_t * m / (sqrt(v) + (double)a.eps_hat);
        d.m[i] = m; d.v[i] = v; d.theta[i] = th;
        if (is_eps) a.eps[0] = (float)th;
        else {
            const int i1 = a.pad_index[i], i2 = a.pad_index2[i];
            const float tf = (float)th;
            a.theta_pad[i1] = tf;
            if (i2 >= 0) a.theta_pad[i2] = tf;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float* mk = a.mirror[k];
                if (mk) { mk[i1] = tf; if (i2 >= 0) mk[i2] = tf; }
            }
        }
    }
    if (a.update && threadIdx.x == 0) a.step[0] = t;
}

cudaError_t hpv_launch_adam(const HpvAdamArgs& a, cudaStream_t s) {
    const int n = a.n_theta + 1;
    int block = ((n + 31) / 32) * 32;
    if (block > 1024) block = 1024;
    hpv_adam_kernel<<<1, block, 0, s>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------
// FP32 FFMA peak probe: every thread runs `iters` rounds of 16 independent fused multiply-adds.
//   variant 0: register operands;  1: multiplier from the constant bank;  2: packed fma.rn.f32x2.
// ---------------------------------------------------------------------------------------------------------
__constant__ float hpv_probe_c[16] = {1.0001f, 0.9999f, 1.0002f, 0.9998f, 1.0003f, 0.9997f, 1.0004f, 0.9996f,
                                      1.0005f, 0.9995f, 1.0006f, 0.9994f, 1.0007f, 0.9993f, 1.0008f, 0.9992f};

__global__ void __launch_bounds__(256) hpv_ffma_probe_kernel(float* out, int iters, int variant) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i) * 1e-3f;
    const float a0 = 1.0f + 1e-7f * (float)threadIdx.x, b0 = 1e-6f;
    if (variant == 0) {
        float a1 = a0 * 0.99999f, a2 = a0 * 1.00001f, a3 = a0 * 0.99998f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    acc[i] = fmaf(acc[i], a0, b0); acc[i + 1] = fmaf(acc[i + 1], a1, b0);
                    acc[i + 2] = fmaf(acc[i + 2], a2, b0); acc[i + 3] = fmaf(acc[i + 3], a3, b0);
                }
            }
        }
    } else if (variant == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], hpv_probe_c[i], b0);
            }
        }
    } else if (variant == 3) {
        // the form the MLP kernels issue: FFMA2 acc, x.F32 (broadcast), UR pair from constant memory, acc
        unsigned long long p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        unsigned long long xa, xb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(xa) : "f"(a0));
        asm("mov.b64 %0, {%1, %1};" : "=l"(xb) : "f"(b0));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    unsigned long long w;
                    const int wi = 2 * ((i + it) & 7);           // uniform, loop-dependent: forces a fresh LDCU
                    asm("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(hpv_probe_c[wi]), "f"(hpv_probe_c[wi + 1]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"((r & 1) ? xa : xb), "l"(w));
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    } else {
        unsigned long long pa, pb, p[8];
        asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a0), "f"(a0 * 0.99999f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b0), "f"(b0));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// flops per launch = grid * block * iters * 64 * 2
cudaError_t hpv_launch_ffma_peak(float* out, int grid, int block, int iters, int variant, cudaStream_t s) {
    hpv_ffma_probe_kernel<<<grid, block, 0, s>>>(out, iters, variant);
    return cudaGetLastError();
}
