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
    return hpv_launch_pdl(hpv_adjproj_kernel, grid, HPV_THREADS, smem, s, a);
}

// TF1 Adam for one parameter (reference order index r; r == n_theta: eps):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  lr_t = lr sqrt(1-b2^t)/(1-b1^t);  theta -= lr_t m/(sqrt(v)+eps_hat)
// (eps_hat outside the bias correction, as tf.train.AdamOptimizer) in float64 master copies; refreshes the fp32
// padded parameters the kernels read -- in global memory and in the constant-memory mirrors this context owns
// (the constant caches are invalidated at the next launch, so the following kernels see the update).
__device__ __forceinline__ void hpv_adam_one(const HpvAdamArgs& a, int r, double g, double lr_t) {
    if (a.grad_out) a.grad_out[r] = g;
    if (!a.update) return;
    const bool is_eps = (r == a.n_theta);
    if (is_eps && !a.train_eps) return;
    const double b1 = a.b1, b2 = a.b2;
    const double m = b1 * a.m[r] + (1.0 - b1) * g;
    const double v = b2 * a.v[r] + (1.0 - b2) * g * g;
    const double th = a.theta[r] - lr_t * m / (sqrt(v) + (double)a.eps_hat);
    a.m[r] = m; a.v[r] = v; a.theta[r] = th;
    if (is_eps) { a.eps[0] = (float)th; return; }
    const int i1 = a.pad_index[r], i2 = a.pad_index2[r];
    const float tf = (float)th;
    a.theta_pad[i1] = tf;
    if (i2 >= 0) a.theta_pad[i2] = tf;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float* mk = a.mirror[k];
        if (mk) { mk[i1] = tf; if (i2 >= 0) mk[i2] = tf; }
    }
}

// The optimizer clock {b1^t, b2^t, t} is read from st_in by everybody and advanced into st_out (the host swaps the
// two buffers between updates) by one thread of the launch: no ordering between CTAs is needed.
__device__ __forceinline__ double hpv_adam_clock(const HpvAdamArgs& a, bool writer) {
    if (!a.update) return 0.0;
    const double p1 = a.st_in[0] * (double)a.b1, p2 = a.st_in[1] * (double)a.b2;
    if (writer) { a.st_out[0] = p1; a.st_out[1] = p2; a.st_out[2] = a.st_in[2] + 1.0; }
    return (double)a.lr * sqrt(1.0 - p2) / (1.0 - p1);
}

__device__ __forceinline__ unsigned long long hpv_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void hpv_st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned hpv_ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float hpv_ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// One warp exchanges chunk `ch` (32 entries, this lane's local value `mine`) with the other ranks and returns the
// sum over the ranks in rank order (see HpvPeerArgs).
__device__ __forceinline__ float hpv_peer_exchange(const HpvPeerArgs& pa, int ch, int lane, float mine) {
    const int par = pa.seq & 1u;
    const size_t voff = ((size_t)par * pa.nranks + pa.rank) * pa.nvp + (size_t)ch * 32 + lane;
    for (int r = 0; r < pa.nranks; ++r) pa.inbox[r][voff] = mine;              // push (remote stores, 128 B per rank)
    __threadfence_system();
    __syncwarp();
    if (lane < pa.nranks)
        hpv_st_release_sys(pa.flags[lane] + ((size_t)par * pa.nranks + pa.rank) * pa.nchunks + ch, pa.seq);
    if (lane < pa.nranks) {                                                     // wait for source rank `lane`
        const unsigned* f = pa.flags[pa.rank] + ((size_t)par * pa.nranks + lane) * pa.nchunks + ch;
        const unsigned long long t0 = hpv_globaltimer();
        while (hpv_ld_acquire_sys(f) != pa.seq) {
            if (hpv_globaltimer() - t0 > pa.timeout_ns) { atomicExch(pa.err, 1u); break; }
        }
    }
    __syncwarp();
    float t = 0.0f;
    const float* in = pa.inbox[pa.rank] + (size_t)par * pa.nranks * pa.nvp + (size_t)ch * 32 + lane;
    for (int r = 0; r < pa.nranks; ++r) t += hpv_ld_relaxed_sys(in + (size_t)r * pa.nvp);
    return t;
}

// CTAs 0 .. nred-1 reduce 32 gradient entries each; with has_loss the last CTA assembles the loss values.  With
// has_peer the chunk is then summed over the GPUs (the losses too), and with has_adam the parameters the chunk's
// entries belong to are updated -- one launch for reduction, exchange and optimizer.
__global__ void __launch_bounds__(1024) hpv_gradreduce_kernel(const HpvGradReduceArgs a, const HpvLossArgs la, int nred,
                                                              const HpvAdamArgs ad, int has_adam,
                                                              const HpvPeerArgs pa, int has_peer, int loss_chunk) {
    __shared__ __align__(16) unsigned char smem[32 * 32 * 4];
    hpv_pdl_wait();          // everything below consumes results of the step's earlier kernels
    if ((int)blockIdx.x >= nred) {
        if (threadIdx.x < 32) {
            hpv_losses_warp(la, threadIdx.x);
            if (has_peer) {
                // la.out[0..7] sits at entry 32 * loss_chunk of the reduce buffer (chunk-aligned)
                __syncwarp();
                const float mine = threadIdx.x < 8 ? la.out[threadIdx.x] : 0.0f;
                const float t = hpv_peer_exchange(pa, loss_chunk, threadIdx.x, mine);
                if (threadIdx.x < 8) la.out[threadIdx.x] = t;
            }
        }
        return;
    }
    HpvCta c;
    c.tid = threadIdx.x; c.nthreads = blockDim.x; c.bid = blockIdx.x; c.nblocks = nred;
    c.smem = smem; c.emu = nullptr;
    float g = hpv_gradreduce_body(c, a);
    if (threadIdx.x < 32) {
        const int i = blockIdx.x * 32 + threadIdx.x;
        if (has_peer) {
            g = hpv_peer_exchange(pa, blockIdx.x, threadIdx.x, i < a.n ? g : 0.0f);
            if (i < a.n) a.grad_pad[i] = g;
        }
        // a timed-out exchange (a peer died) leaves stale inbox data in g: keep the parameters as they are; the host
        // reads the error word after its next synchronisation (check_peer_error) and fails the call
        const bool exchange_failed = has_peer && *reinterpret_cast<volatile unsigned*>(pa.err) != 0u;
        if (has_adam && !exchange_failed) {
            const double lr_t = hpv_adam_clock(ad, i == 0);
            if (i < a.n) {
                const int r = ad.ref_index[i];
                if (r >= 0) hpv_adam_one(ad, r, (double)g, lr_t);
            }
        }
    }
}

cudaError_t hpv_launch_gradreduce(const HpvGradReduceArgs& a, const HpvLossArgs* la, const HpvAdamArgs* adam,
                                  const HpvPeerArgs* peer, int loss_off, cudaStream_t s) {
    const int nred = (a.n + 31) / 32;
    HpvLossArgs l0;
    memset(&l0, 0, sizeof(l0));
    HpvAdamArgs a0;
    memset(&a0, 0, sizeof(a0));
    HpvPeerArgs p0;
    memset(&p0, 0, sizeof(p0));
    // 32 groups of partials per CTA when there are many of them (the sum is latency-bound otherwise).  With the peer
    // exchange every CTA spins on flags that the matching CTA of the other ranks sets: all CTAs of the grid must be
    // co-resident (forward progress must not depend on the dispatch order), which 1024-thread CTAs (2 per SM)
    // only guarantee up to 2 x 148 of them -- beyond that the grid runs 256-thread CTAs (8 per SM).
    int block = a.n_parts >= 128 ? 1024 : 256;
    if (peer && nred + 1 > 256) block = 256;
    return hpv_launch_pdl(hpv_gradreduce_kernel, nred + (la ? 1 : 0), block, 0, s, a, la ? *la : l0, nred, adam ? *adam : a0,
                          adam ? 1 : 0, peer ? *peer : p0, peer ? 1 : 0, loss_off / 32);
}

// Stand-alone form (after the NCCL all-reduce of the multi-GPU step, and to un-pad a gradient for the host):
// one CTA, thread-strided over the parameters (reference order, eps last).
__global__ void __launch_bounds__(1024) hpv_adam_kernel(const HpvAdamArgs a) {
    if (a.losses_in && a.grad_out && threadIdx.x < 8) a.grad_out[a.n_theta + 1 + threadIdx.x] = (double)a.losses_in[threadIdx.x];
    const double lr_t = hpv_adam_clock(a, threadIdx.x == 0);
    for (int r = threadIdx.x; r <= a.n_theta; r += blockDim.x) {
        const double g = (double)(r == a.n_theta ? a.grad_pad[a.theta_pad_n] : a.grad_pad[a.pad_index[r]]);
        hpv_adam_one(a, r, g, lr_t);
    }
}

cudaError_t hpv_launch_adam(const HpvAdamArgs& a, cudaStream_t s) {
    const int n = a.n_theta + 1;
    int block = ((n + 31) / 32) * 32;
    if (block > 1024) block = 1024;
    hpv_adam_kernel<<<1, block, 0, s>>>(a);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) hpv_param_scatter_kernel(const HpvParamScatterArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.theta_pad_n) {
        const float v = a.blob_pad[i];
        a.theta_pad[i] = v;
#pragma unroll
        for (int k = 0; k < 3; ++k) if (a.mirror[k]) a.mirror[k][i] = v;
    }
    if (i == 0) a.eps[0] = a.blob_pad[a.theta_pad_n];
    if (i < a.n_master) a.master[i] = a.blob_master[i];
}

cudaError_t hpv_launch_param_scatter(const HpvParamScatterArgs& a, cudaStream_t s) {
    const int n = a.theta_pad_n > a.n_master ? a.theta_pad_n : a.n_master;
    hpv_param_scatter_kernel<<<(n + 255) / 256, 256, 0, s>>>(a);
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
