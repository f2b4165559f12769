// Development probe (not part of the product): the tcgen05 / TMEM building blocks of the tensor-core layer product,
// each checked against a float64 host result, and the MLP phase of the forward kernel in its tensor-core form
// timed against the point count.   Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//   umma_probe.bin TEST      TEST = 1 SS K-major | 2 TS (A in TMEM) | 3 SS with N = 24 at M = 128 |
//                                   4 MN-major operands, M = 64 (weight-gradient shape) | 5 3xTF32 accuracy (TS) |
//                                   6 MLP phase (value + two tangents, [2,20,20,20,1]) accuracy and time
// Every test runs in its own process (a faulting kernel poisons the context).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../../hp-vpinns_b200/csrc/hpv_umma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// ------------------------------------------------------------------------------------------------------------------
// D[128][N] = A[128][K] . B[N][K]^T, one CTA of 128 threads.  mode 0: A from shared memory, 1: A from TMEM.
// split 0: operands used as they are (TF32 truncation by the tensor core), 1: 3xTF32.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_gemm(const float* A, const float* B, float* D, int N, int K, int mode, int split) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t* sAhi = reinterpret_cast<uint32_t*>(sm);
    uint32_t* sAlo = sAhi + 128 * K;
    uint32_t* sBhi = sAlo + 128 * K;
    uint32_t* sBlo = sBhi + N * K;
    if (warp == 0) hpv_tmem_alloc(&tbase, 256);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i % K;
        uint32_t hi, lo;
        if (split) hpv_split_tf32(B[i], hi, lo); else { hi = __float_as_uint(B[i]); lo = 0; }
        const int w = (k / 4) * (N * 4) + n * 4 + (k % 4);
        sBhi[w] = hi; sBlo[w] = lo;
    }
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t colAhi = 32, colAlo = 32 + 64;        // D: [0, 32)
    for (int k4 = 0; k4 < K / 4; ++k4) {
        uint32_t hi[4], lo[4];
        for (int j = 0; j < 4; ++j) {
            const float v = A[tid * K + 4 * k4 + j];
            if (split) hpv_split_tf32(v, hi[j], lo[j]); else { hi[j] = __float_as_uint(v); lo[j] = 0; }
        }
        if (mode == 0) {
            const int w = k4 * (128 * 4) + tid * 4;
            for (int j = 0; j < 4; ++j) { sAhi[w + j] = hi[j]; sAlo[w + j] = lo[j]; }
        } else {
            hpv_tmem_st4(tb + lane_base + colAhi + 4 * k4, hi[0], hi[1], hi[2], hi[3]);
            hpv_tmem_st4(tb + lane_base + colAlo + 4 * k4, lo[0], lo[1], lo[2], lo[3]);
        }
    }
    if (mode == 1) hpv_tmem_wait_st();
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        hpv_tc_fence_after();
        const uint32_t idesc = hpv_umma_idesc_tf32(128, N, 0, 0);
        uint32_t acc = 0;
        for (int term = split ? 0 : 2; term < 3; ++term) {           // (Alo,Bhi), (Ahi,Blo), (Ahi,Bhi)
            const uint32_t* sA = term == 0 ? sAlo : sAhi;
            const uint32_t* sB = term == 1 ? sBlo : sBhi;
            const uint32_t cA = term == 0 ? colAlo : colAhi;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t bd = hpv_umma_desc(hpv_smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
                if (mode == 0) {
                    const uint64_t ad = hpv_umma_desc(hpv_smem_u32(sA) + ks * 2 * (128 * 16), 128 * 16, 128);
                    hpv_umma_ss(tb, ad, bd, idesc, acc);
                } else {
                    hpv_umma_ts(tb, tb + cA + ks * 8, bd, idesc, acc);
                }
                acc = 1;
            }
        }
        hpv_umma_commit(&bar);
    }
    hpv_mbar_wait(&bar, 0);
    hpv_tc_fence_after();
    for (int c4 = 0; c4 < N / 4; ++c4) {
        uint32_t r[4];
        hpv_tmem_ld4(tb + lane_base + 4 * c4, r);
        hpv_tmem_wait_ld();
        for (int j = 0; j < 4; ++j) D[tid * N + 4 * c4 + j] = __uint_as_float(r[j]);
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 256);
}

// ------------------------------------------------------------------------------------------------------------------
// D[64][N] = sum_k At[k][m] * Bt[k][n], K = 128: both operands MN-major in shared memory (the weight-gradient shape:
// At = activations [point][unit], Bt = adjoints [point][unit]).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_gemm_mn(const float* At, const float* Bt, float* D, int N, int K) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = 64;
    uint32_t* sA = reinterpret_cast<uint32_t*>(sm);          // [M/4][K/8][8][4]
    uint32_t* sB = sA + M * K;                                 // [N/4][K/8][8][4]
    if (warp == 0) hpv_tmem_alloc(&tbase, 32);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    const int SBO = (K / 8) * 128;                             // bytes between blocks of 4 rows; LBO = 128
    for (int i = tid; i < M * K; i += 128) {
        const int k = i / M, m = i % M;
        sA[((m / 4) * SBO + (k / 8) * 128 + (k % 8) * 16 + (m % 4) * 4) / 4] = __float_as_uint(At[i]);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int k = i / N, n = i % N;
        sB[((n / 4) * SBO + (k / 8) * 128 + (k % 8) * 16 + (n % 4) * 4) / 4] = __float_as_uint(Bt[i]);
    }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase;
    if (tid == 0) {
        const uint32_t idesc = hpv_umma_idesc_tf32(64, N, 1, 1);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ad = hpv_umma_desc(hpv_smem_u32(sA) + ks * 128, 128, SBO);
            const uint64_t bd = hpv_umma_desc(hpv_smem_u32(sB) + ks * 128, 128, SBO);
            hpv_umma_ss(tb, ad, bd, idesc, ks > 0);
        }
        hpv_umma_commit(&bar);
    }
    hpv_mbar_wait(&bar, 0);
    hpv_tc_fence_after();
    // M = 64: row r lives in lane (r % 16) + 32 * (r / 16)
    for (int c4 = 0; c4 < N / 4; ++c4) {
        uint32_t r[4];
        hpv_tmem_ld4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, r);
        hpv_tmem_wait_ld();
        if (lane < 16)
            for (int j = 0; j < 4; ++j) D[(warp * 16 + lane) * N + 4 * c4 + j] = __uint_as_float(r[j]);
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 32);
}

// ------------------------------------------------------------------------------------------------------------------
// MLP phase, tensor-core form: net [2, 20, 20, 20, 1], tanh, channels (value, d/dx, d/dy).
// CTA = 256 threads = 8 warps; warp w owns TMEM sub-partition w % 4 (32 points of the 128-point tile) and the
// units [10 (w / 4), 10 (w / 4) + 10).  Activations of a layer are split hi/lo and stored to TMEM as the A operand
// of the next product; the weights (+ bias row) sit in shared memory as K-major B tiles; z comes back from TMEM.
// ------------------------------------------------------------------------------------------------------------------
#define HP 20
#define KP 24
#define NPAD 32
#define NCH 3
#define NHH 2                                     // hidden-to-hidden products
struct MlpTheta { float W1[2][HP], b1[HP], Wh[NHH][HP][HP], bh[NHH][HP], Wo[HP], bo; };

__device__ __forceinline__ float tanh_acc(float z) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}

__global__ void __launch_bounds__(256, 2) k_mlp(const MlpTheta* __restrict__ thg, float* __restrict__ out, int n_tiles, int split,
                                                 long long* cyc_out) {
    __shared__ __align__(128) uint32_t sB[NHH][2][KP / 4][NPAD][4];     // [layer][hi|lo][k chunk][n][4]
    __shared__ MlpTheta th;
    __shared__ __align__(8) uint64_t bar[NCH];
    __shared__ uint32_t tbase;
    __shared__ float s_part[NCH][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = warp & 3, half = warp >> 2, u0 = half * (HP / 2);
    const int prow = sub * 32 + lane;
    if (warp == 0) hpv_tmem_alloc(&tbase, 256);
    if (tid == 0) { for (int c = 0; c < NCH; ++c) hpv_mbar_init(&bar[c], 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < (int)(sizeof(MlpTheta) / 4); i += 256) reinterpret_cast<float*>(&th)[i] = reinterpret_cast<const float*>(thg)[i];
    for (int i = tid; i < NHH * 2 * (KP / 4) * NPAD * 4; i += 256) (&sB[0][0][0][0][0])[i] = 0u;
    __syncthreads();
    for (int i = tid; i < NHH * NPAD * KP; i += 256) {
        const int l = i / (NPAD * KP), n = (i / KP) % NPAD, k = i % KP;
        float v = 0.0f;
        if (n < HP) v = k < HP ? th.Wh[l][k][n] : (k == HP ? th.bh[l][n] : 0.0f);
        uint32_t hi, lo;
        if (split) hpv_split_tf32(v, hi, lo); else { hi = __float_as_uint(v); lo = 0; }
        sB[l][0][k / 4][n][k % 4] = hi; sB[l][1][k / 4][n][k % 4] = lo;
    }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase, lane_base = (uint32_t)(sub * 32) << 16;
    const uint32_t colD = 0, colAhi = NCH * NPAD, colAlo = colAhi + NCH * KP;
    // constant columns of the A operands: unit HP of the value channel is the bias input 1, the rest of the padding 0
    if (half == 0) {
        for (int c = 0; c < NCH; ++c) {
            hpv_tmem_st4(tb + lane_base + colAhi + c * KP + HP, c == 0 ? __float_as_uint(1.0f) : 0u, 0u, 0u, 0u);
            hpv_tmem_st4(tb + lane_base + colAlo + c * KP + HP, 0u, 0u, 0u, 0u);
        }
        hpv_tmem_wait_st();
    }
    const uint32_t idesc = hpv_umma_idesc_tf32(128, NPAD, 0, 0);
    uint32_t phase = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int p = tile * 128 + prow;
        const float x = -1.0f + 2.0f * (float)(p % 1024) / 1024.0f, y = -1.0f + 2.0f * (float)(p / 1024 % 1024) / 1024.0f;
        float zv[HP / 2], zx[HP / 2], zy[HP / 2];
#pragma unroll
        for (int j = 0; j < HP / 2; ++j) {
            const int u = u0 + j;
            zx[j] = th.W1[0][u]; zy[j] = th.W1[1][u];
            zv[j] = fmaf(y, zy[j], fmaf(x, zx[j], th.b1[u]));
        }
#pragma unroll 1
        for (int l = 0; l <= NHH; ++l) {
            if (l > 0) {
                // z of hidden layer l+1 from TMEM, channel by channel as the products complete
                uint32_t r8[8], r2[2];
                hpv_mbar_wait(&bar[0], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 0 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 0 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zv[j] = __uint_as_float(r8[j]);
                zv[8] = __uint_as_float(r2[0]); zv[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[1], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 1 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 1 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zx[j] = __uint_as_float(r8[j]);
                zx[8] = __uint_as_float(r2[0]); zx[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[2], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 2 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 2 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zy[j] = __uint_as_float(r8[j]);
                zy[8] = __uint_as_float(r2[0]); zy[9] = __uint_as_float(r2[1]);
                phase ^= 1;
            }
            float hv[HP / 2], hx[HP / 2], hy[HP / 2];
#pragma unroll
            for (int j = 0; j < HP / 2; ++j) {
                const float a = tanh_acc(zv[j]), s1 = fmaf(-a, a, 1.0f);
                hv[j] = a; hx[j] = s1 * zx[j]; hy[j] = s1 * zy[j];
            }
            if (l == NHH) {
                float pv = 0.0f, px = 0.0f, py = 0.0f;
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) { const float w = th.Wo[u0 + j]; pv = fmaf(hv[j], w, pv); px = fmaf(hx[j], w, px); py = fmaf(hy[j], w, py); }
                if (half == 1) { s_part[0][prow] = pv; s_part[1][prow] = px; s_part[2][prow] = py; }
                __syncthreads();
                if (half == 0) {
                    out[(size_t)p * 3 + 0] = pv + s_part[0][prow] + th.bo;
                    out[(size_t)p * 3 + 1] = px + s_part[1][prow];
                    out[(size_t)p * 3 + 2] = py + s_part[2][prow];
                }
                break;
            }
            // split and store the A operand of the next product
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float* h = c == 0 ? hv : (c == 1 ? hx : hy);
                uint32_t hi[HP / 2], lo[HP / 2];
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) {
                    if (split) hpv_split_tf32(h[j], hi[j], lo[j]); else { hi[j] = __float_as_uint(h[j]); lo[j] = 0; }
                }
                const uint32_t a_hi = tb + lane_base + colAhi + c * KP + u0, a_lo = tb + lane_base + colAlo + c * KP + u0;
                uint32_t t8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t8[j] = hi[j];
                hpv_tmem_st8(a_hi, t8); hpv_tmem_st2(a_hi + 8, hi[8], hi[9]);
                if (split) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) t8[j] = lo[j];
                    hpv_tmem_st8(a_lo, t8); hpv_tmem_st2(a_lo + 8, lo[8], lo[9]);
                }
            }
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                hpv_tc_fence_after();
                for (int c = 0; c < NCH; ++c) {
                    uint32_t acc = 0;
                    for (int term = split ? 0 : 2; term < 3; ++term) {
                        const uint32_t cA = (term == 0 ? colAlo : colAhi) + c * KP;
                        const uint32_t bB = hpv_smem_u32(&sB[l][term == 1 ? 1 : 0][0][0][0]);
                        for (int ks = 0; ks < KP / 8; ++ks) {
                            hpv_umma_ts(tb + colD + c * NPAD, tb + cA + ks * 8, hpv_umma_desc(bB + ks * 2 * (NPAD * 16), NPAD * 16, 128), idesc, acc);
                            acc = 1;
                        }
                    }
                    hpv_umma_commit(&bar[c]);
                }
            }
        }
    }
    if (tid == 0 && cyc_out) cyc_out[blockIdx.x] = clock64() - t_start;
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 256);
}


// ------------------------------------------------------------------------------------------------------------------
// M = 64 layout finder: D[64][N] = A[64][K] . B[N][K]^T with K-major operands (known good), all 128 lanes dumped.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) k_gemm_m64(const float* A, const float* B, float* Dall, int N, int K) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t* sA = reinterpret_cast<uint32_t*>(sm);
    uint32_t* sB = sA + 64 * K;
    if (warp == 0) hpv_tmem_alloc(&tbase, 32);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < 64 * K; i += 128) { const int m = i / K, k = i % K; sA[(k / 4) * (64 * 4) + m * 4 + (k % 4)] = __float_as_uint(A[i]); }
    for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i % K; sB[(k / 4) * (N * 4) + n * 4 + (k % 4)] = __float_as_uint(B[i]); }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase;
    // clear D first (all 128 lanes) so that untouched lanes read as a marker
    for (int c4 = 0; c4 < N / 4; ++c4) hpv_tmem_st4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u);
    hpv_tmem_wait_st();
    hpv_tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        hpv_tc_fence_after();
        const uint32_t idesc = hpv_umma_idesc_tf32(64, N, 0, 0);
        for (int ks = 0; ks < K / 8; ++ks)
            hpv_umma_ss(tb, hpv_umma_desc(hpv_smem_u32(sA) + ks * 2 * (64 * 16), 64 * 16, 128), hpv_umma_desc(hpv_smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128), idesc, ks > 0);
        hpv_umma_commit(&bar);
    }
    hpv_mbar_wait(&bar, 0);
    hpv_tc_fence_after();
    for (int c4 = 0; c4 < N / 4; ++c4) {
        uint32_t r[4];
        hpv_tmem_ld4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, r);
        hpv_tmem_wait_ld();
        for (int j = 0; j < 4; ++j) Dall[tid * N + 4 * c4 + j] = __uint_as_float(r[j]);
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 32);
}

// MN-major operands, generic physical strides and descriptor fields; all 128 lanes dumped.
//   byte(row r, k) = (r%4)*4 + (k%8)*16 + (r/4)*strideR4 + (k/8)*strideK8
__global__ void __launch_bounds__(128, 1) k_gemm_mn2(const float* At, const float* Bt, float* Dall, int N, int K, int strideR4, int strideK8,
                                                      int lbo, int sbo, int kstep) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = 64;
    uint32_t* sA = reinterpret_cast<uint32_t*>(sm);
    uint32_t* sB = sA + M * K;
    if (warp == 0) hpv_tmem_alloc(&tbase, 32);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < M * K; i += 128) { const int k = i / M, m = i % M; sA[((m / 4) * strideR4 + (k / 8) * strideK8 + (k % 8) * 16 + (m % 4) * 4) / 4] = __float_as_uint(At[i]); }
    const int strideR4B = strideR4 == 128 ? 128 : strideR4, strideK8B = strideK8 == 128 ? 128 : (N / 4) * 128;
    for (int i = tid; i < N * K; i += 128) { const int k = i / N, n = i % N; sB[((n / 4) * strideR4B + (k / 8) * strideK8B + (k % 8) * 16 + (n % 4) * 4) / 4] = __float_as_uint(Bt[i]); }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase;
    for (int c4 = 0; c4 < N / 4; ++c4) hpv_tmem_st4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u);
    hpv_tmem_wait_st();
    hpv_tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        hpv_tc_fence_after();
        const uint32_t idesc = hpv_umma_idesc_tf32(64, N, 1, 1);
        // descriptor fields: (lbo, sbo) as given for A; for B the same roles with B's strides
        const int lboB = lbo == strideK8 ? strideK8B : strideR4B, sboB = sbo == strideK8 ? strideK8B : strideR4B;
        for (int ks = 0; ks < K / 8; ++ks)
            hpv_umma_ss(tb, hpv_umma_desc(hpv_smem_u32(sA) + ks * strideK8, lbo, sbo), hpv_umma_desc(hpv_smem_u32(sB) + ks * strideK8B, lboB, sboB), idesc, ks > 0);
        (void)kstep;
        hpv_umma_commit(&bar);
    }
    hpv_mbar_wait(&bar, 0);
    hpv_tc_fence_after();
    for (int c4 = 0; c4 < N / 4; ++c4) {
        uint32_t r[4];
        hpv_tmem_ld4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, r);
        hpv_tmem_wait_ld();
        for (int j = 0; j < 4; ++j) Dall[tid * N + 4 * c4 + j] = __uint_as_float(r[j]);
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 32);
}


// MN-major TF32 operands: the dedicated "128-byte swizzle with 32-byte base" layout (cute::UMMA::Layout_MN_SW128_32B_Atom,
// LayoutType 1): atoms of 4 k x 32 rows = 512 bytes,
//   byte(row r, k) = (k/4)*SBO + (r/32)*LBO + (k%4)*128 + ((((r%32)/8) ^ (k%4)) * 32) + (r%8)*4
__global__ void __launch_bounds__(128, 1) k_gemm_mn3(const float* At, const float* Bt, float* Dall, int N, int K, int variant) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = 64;
    unsigned char* sA = sm;
    const int nbA = M / 32, nbB = (N + 31) / 32;
    const int sboA = nbA * 512, sboB = nbB * 512;              // k-block (4 k) stride; row blocks adjacent (LBO = 512)
    unsigned char* sB = sm + (K / 4) * sboA;
    if (warp == 0) hpv_tmem_alloc(&tbase, 32);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < ((K / 4) * (sboA + sboB)) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    __syncthreads();
    for (int i = tid; i < M * K; i += 128) {
        const int k = i / M, m = i % M;
        *reinterpret_cast<float*>(sA + (k / 4) * sboA + (m / 32) * 512 + (k % 4) * 128 + ((((m % 32) / 8) ^ (k % 4)) * 32) + (m % 8) * 4) = At[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        const int k = i / N, n = i % N;
        *reinterpret_cast<float*>(sB + (k / 4) * sboB + (n / 32) * 512 + (k % 4) * 128 + ((((n % 32) / 8) ^ (k % 4)) * 32) + (n % 8) * 4) = Bt[i];
    }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase;
    for (int c4 = 0; c4 < N / 4; ++c4) hpv_tmem_st4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u, 0x7fc00000u);
    hpv_tmem_wait_st();
    hpv_tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        hpv_tc_fence_after();
        const uint32_t idesc = hpv_umma_idesc_tf32(64, N, 1, 1);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t lboA = variant == 0 ? 512 : sboA, sbA = variant == 0 ? sboA : 512;
            const uint32_t lboB = variant == 0 ? 512 : sboB, sbB = variant == 0 ? sboB : 512;
            hpv_umma_ss(tb, hpv_umma_desc(hpv_smem_u32(sA) + ks * 2 * sboA, lboA, sbA, 1), hpv_umma_desc(hpv_smem_u32(sB) + ks * 2 * sboB, lboB, sbB, 1), idesc, ks > 0);
        }
        hpv_umma_commit(&bar);
    }
    hpv_mbar_wait(&bar, 0);
    hpv_tc_fence_after();
    for (int c4 = 0; c4 < N / 4; ++c4) {
        uint32_t r[4];
        hpv_tmem_ld4(tb + ((uint32_t)(warp * 32) << 16) + 4 * c4, r);
        hpv_tmem_wait_ld();
        for (int j = 0; j < 4; ++j) Dall[tid * N + 4 * c4 + j] = __uint_as_float(r[j]);
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 32);
}

__device__ __forceinline__ bool hpv_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}


// MLP phase, second form: truncation split (hi = x & 0xffffe000, lo = x - hi), the whole warp 0 runs the issue loop
// with the descriptors hoisted (uniform registers) and one elected lane issuing, MMAs interleaved over the channels
// (ORDER 1) so that consecutive instructions accumulate into different D tiles; ND = N of the instruction (32 or 24).
template <int ORDER, int ND>
__global__ void __launch_bounds__(256, 2) k_mlp2(const MlpTheta* __restrict__ thg, float* __restrict__ out, int n_tiles, long long* cyc_out) {
    __shared__ __align__(128) uint32_t sB[NHH][2][KP / 4][NPAD][4];
    __shared__ MlpTheta th;
    __shared__ __align__(8) uint64_t bar[NCH];
    __shared__ uint32_t tbase;
    __shared__ float s_part[NCH][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = warp & 3, half = warp >> 2, u0 = half * (HP / 2);
    const int prow = sub * 32 + lane;
    if (warp == 0) hpv_tmem_alloc(&tbase, 256);
    if (tid == 0) { for (int c = 0; c < NCH; ++c) hpv_mbar_init(&bar[c], 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < (int)(sizeof(MlpTheta) / 4); i += 256) reinterpret_cast<float*>(&th)[i] = reinterpret_cast<const float*>(thg)[i];
    for (int i = tid; i < NHH * 2 * (KP / 4) * NPAD * 4; i += 256) (&sB[0][0][0][0][0])[i] = 0u;
    __syncthreads();
    for (int i = tid; i < NHH * NPAD * KP; i += 256) {
        const int l = i / (NPAD * KP), n = (i / KP) % NPAD, k = i % KP;
        float v = 0.0f;
        if (n < HP) v = k < HP ? th.Wh[l][k][n] : (k == HP ? th.bh[l][n] : 0.0f);
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        sB[l][0][k / 4][n][k % 4] = hi; sB[l][1][k / 4][n][k % 4] = __float_as_uint(v - __uint_as_float(hi));
    }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0), lane_base = (uint32_t)(sub * 32) << 16;
    const uint32_t colD = 0, colAhi = NCH * NPAD, colAlo = colAhi + NCH * KP;
    if (half == 0) {
        for (int c = 0; c < NCH; ++c) {
            hpv_tmem_st4(tb + lane_base + colAhi + c * KP + HP, c == 0 ? __float_as_uint(1.0f) : 0u, 0u, 0u, 0u);
            hpv_tmem_st4(tb + lane_base + colAlo + c * KP + HP, 0u, 0u, 0u, 0u);
        }
        hpv_tmem_wait_st();
    }
    const uint32_t idesc = hpv_umma_idesc_tf32(128, ND, 0, 0);
    uint32_t phase = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int p = tile * 128 + prow;
        const float x = -1.0f + 2.0f * (float)(p % 1024) / 1024.0f, y = -1.0f + 2.0f * (float)(p / 1024 % 1024) / 1024.0f;
        float zv[HP / 2], zx[HP / 2], zy[HP / 2];
#pragma unroll
        for (int j = 0; j < HP / 2; ++j) {
            const int u = u0 + j;
            zx[j] = th.W1[0][u]; zy[j] = th.W1[1][u];
            zv[j] = fmaf(y, zy[j], fmaf(x, zx[j], th.b1[u]));
        }
#pragma unroll 1
        for (int l = 0; l <= NHH; ++l) {
            if (l > 0) {
                uint32_t r8[8], r2[2];
                hpv_mbar_wait(&bar[0], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 0 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 0 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zv[j] = __uint_as_float(r8[j]);
                zv[8] = __uint_as_float(r2[0]); zv[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[1], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 1 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 1 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zx[j] = __uint_as_float(r8[j]);
                zx[8] = __uint_as_float(r2[0]); zx[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[2], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 2 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 2 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zy[j] = __uint_as_float(r8[j]);
                zy[8] = __uint_as_float(r2[0]); zy[9] = __uint_as_float(r2[1]);
                phase ^= 1;
            }
            float hv[HP / 2], hx[HP / 2], hy[HP / 2];
#pragma unroll
            for (int j = 0; j < HP / 2; ++j) {
                const float a = tanh_acc(zv[j]), s1 = fmaf(-a, a, 1.0f);
                hv[j] = a; hx[j] = s1 * zx[j]; hy[j] = s1 * zy[j];
            }
            if (l == NHH) {
                float pv = 0.0f, px = 0.0f, py = 0.0f;
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) { const float w = th.Wo[u0 + j]; pv = fmaf(hv[j], w, pv); px = fmaf(hx[j], w, px); py = fmaf(hy[j], w, py); }
                if (half == 1) { s_part[0][prow] = pv; s_part[1][prow] = px; s_part[2][prow] = py; }
                __syncthreads();
                if (half == 0) {
                    out[(size_t)p * 3 + 0] = pv + s_part[0][prow] + th.bo;
                    out[(size_t)p * 3 + 1] = px + s_part[1][prow];
                    out[(size_t)p * 3 + 2] = py + s_part[2][prow];
                }
                break;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float* h = c == 0 ? hv : (c == 1 ? hx : hy);
                uint32_t hi[HP / 2], lo[HP / 2];
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) { hi[j] = __float_as_uint(h[j]) & 0xffffe000u; lo[j] = __float_as_uint(h[j] - __uint_as_float(hi[j])); }
                const uint32_t a_hi = tb + lane_base + colAhi + c * KP + u0, a_lo = tb + lane_base + colAlo + c * KP + u0;
                uint32_t t8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t8[j] = hi[j];
                hpv_tmem_st8(a_hi, t8); hpv_tmem_st2(a_hi + 8, hi[8], hi[9]);
#pragma unroll
                for (int j = 0; j < 8; ++j) t8[j] = lo[j];
                hpv_tmem_st8(a_lo, t8); hpv_tmem_st2(a_lo + 8, lo[8], lo[9]);
            }
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                hpv_tc_fence_after();
                const uint32_t bhi = hpv_smem_u32(&sB[l][0][0][0][0]), blo = hpv_smem_u32(&sB[l][1][0][0][0]);
                if (ORDER == 0) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
#pragma unroll
                            for (int ks = 0; ks < KP / 8; ++ks) {
                                const uint32_t a = tb + (term == 0 ? colAlo : colAhi) + c * KP + ks * 8;
                                const uint64_t bd = hpv_umma_desc((term == 1 ? blo : bhi) + ks * 2 * (NPAD * 16), NPAD * 16, 128);
                                if (hpv_elect_one()) hpv_umma_ts(tb + colD + c * NPAD, a, bd, idesc, (term | ks) != 0);
                            }
                        }
                        if (hpv_elect_one()) hpv_umma_commit(&bar[c]);
                    }
                } else {
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
#pragma unroll
                        for (int ks = 0; ks < KP / 8; ++ks) {
                            const uint64_t bd = hpv_umma_desc((term == 1 ? blo : bhi) + ks * 2 * (NPAD * 16), NPAD * 16, 128);
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const uint32_t a = tb + (term == 0 ? colAlo : colAhi) + c * KP + ks * 8;
                                if (hpv_elect_one()) hpv_umma_ts(tb + colD + c * NPAD, a, bd, idesc, (term | ks) != 0);
                                if (term == 2 && ks == KP / 8 - 1) { if (hpv_elect_one()) hpv_umma_commit(&bar[c]); }
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    if (tid == 0 && cyc_out) cyc_out[blockIdx.x] = clock64() - t_start;
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 256);
}


// MLP phase, third form: as k_mlp2 (interleaved order) but the issue block is specialised on the TMEM base address
// (0 or 256: two CTAs per SM, 256 columns each), so that every operand of the tensor-core instructions is a
// compile-time constant or uniform arithmetic and the elected thread issues them back to back (no R2UR chains).
template <uint32_t TB, int ND>
__device__ __forceinline__ void mlp3_issue(uint32_t bhi, uint32_t blo, uint64_t* bar) {
    constexpr uint32_t colD = 0, colAhi = NCH * NPAD, colAlo = colAhi + NCH * KP;
    constexpr uint32_t idesc = hpv_umma_idesc_tf32(128, ND, 0, 0);
#pragma unroll
    for (int term = 0; term < 3; ++term) {
#pragma unroll
        for (int ks = 0; ks < KP / 8; ++ks) {
            const uint64_t bd = hpv_umma_desc((term == 1 ? blo : bhi) + ks * 2 * (NPAD * 16), NPAD * 16, 128);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                hpv_umma_ts(TB + colD + c * NPAD, TB + (term == 0 ? colAlo : colAhi) + c * KP + ks * 8, bd, idesc, (term | ks) != 0);
                if (term == 2 && ks == KP / 8 - 1) hpv_umma_commit(&bar[c]);
            }
        }
    }
}

template <int ND>
__global__ void __launch_bounds__(256, 2) k_mlp3(const MlpTheta* __restrict__ thg, float* __restrict__ out, int n_tiles, long long* cyc_out) {
    __shared__ __align__(128) uint32_t sB[NHH][2][KP / 4][NPAD][4];
    __shared__ MlpTheta th;
    __shared__ __align__(8) uint64_t bar[NCH];
    __shared__ uint32_t tbase;
    __shared__ float s_part[NCH][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = warp & 3, half = warp >> 2, u0 = half * (HP / 2);
    const int prow = sub * 32 + lane;
    if (warp == 0) hpv_tmem_alloc(&tbase, 256);
    if (tid == 0) { for (int c = 0; c < NCH; ++c) hpv_mbar_init(&bar[c], 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < (int)(sizeof(MlpTheta) / 4); i += 256) reinterpret_cast<float*>(&th)[i] = reinterpret_cast<const float*>(thg)[i];
    for (int i = tid; i < NHH * 2 * (KP / 4) * NPAD * 4; i += 256) (&sB[0][0][0][0][0])[i] = 0u;
    __syncthreads();
    for (int i = tid; i < NHH * NPAD * KP; i += 256) {
        const int l = i / (NPAD * KP), n = (i / KP) % NPAD, k = i % KP;
        float v = 0.0f;
        if (n < HP) v = k < HP ? th.Wh[l][k][n] : (k == HP ? th.bh[l][n] : 0.0f);
        const uint32_t hi = __float_as_uint(v) & 0xffffe000u;
        sB[l][0][k / 4][n][k % 4] = hi; sB[l][1][k / 4][n][k % 4] = __float_as_uint(v - __uint_as_float(hi));
    }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = tbase, lane_base = (uint32_t)(sub * 32) << 16;
    if (tb != 0u && tb != 256u) { asm volatile("trap;"); }
    const uint32_t colD = 0, colAhi = NCH * NPAD, colAlo = colAhi + NCH * KP;
    if (half == 0) {
        for (int c = 0; c < NCH; ++c) {
            hpv_tmem_st4(tb + lane_base + colAhi + c * KP + HP, c == 0 ? __float_as_uint(1.0f) : 0u, 0u, 0u, 0u);
            hpv_tmem_st4(tb + lane_base + colAlo + c * KP + HP, 0u, 0u, 0u, 0u);
        }
        hpv_tmem_wait_st();
    }
    uint32_t phase = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int p = tile * 128 + prow;
        const float x = -1.0f + 2.0f * (float)(p % 1024) / 1024.0f, y = -1.0f + 2.0f * (float)(p / 1024 % 1024) / 1024.0f;
        float zv[HP / 2], zx[HP / 2], zy[HP / 2];
#pragma unroll
        for (int j = 0; j < HP / 2; ++j) {
            const int u = u0 + j;
            zx[j] = th.W1[0][u]; zy[j] = th.W1[1][u];
            zv[j] = fmaf(y, zy[j], fmaf(x, zx[j], th.b1[u]));
        }
#pragma unroll 1
        for (int l = 0; l <= NHH; ++l) {
            if (l > 0) {
                uint32_t r8[8], r2[2];
                hpv_mbar_wait(&bar[0], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 0 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 0 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zv[j] = __uint_as_float(r8[j]);
                zv[8] = __uint_as_float(r2[0]); zv[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[1], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 1 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 1 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zx[j] = __uint_as_float(r8[j]);
                zx[8] = __uint_as_float(r2[0]); zx[9] = __uint_as_float(r2[1]);
                hpv_mbar_wait(&bar[2], phase); hpv_tc_fence_after();
                hpv_tmem_ld8(tb + lane_base + colD + 2 * NPAD + u0, r8); hpv_tmem_ld2(tb + lane_base + colD + 2 * NPAD + u0 + 8, r2);
                hpv_tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) zy[j] = __uint_as_float(r8[j]);
                zy[8] = __uint_as_float(r2[0]); zy[9] = __uint_as_float(r2[1]);
                phase ^= 1;
            }
            float hv[HP / 2], hx[HP / 2], hy[HP / 2];
#pragma unroll
            for (int j = 0; j < HP / 2; ++j) {
                const float a = tanh_acc(zv[j]), s1 = fmaf(-a, a, 1.0f);
                hv[j] = a; hx[j] = s1 * zx[j]; hy[j] = s1 * zy[j];
            }
            if (l == NHH) {
                float pv = 0.0f, px = 0.0f, py = 0.0f;
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) { const float w = th.Wo[u0 + j]; pv = fmaf(hv[j], w, pv); px = fmaf(hx[j], w, px); py = fmaf(hy[j], w, py); }
                if (half == 1) { s_part[0][prow] = pv; s_part[1][prow] = px; s_part[2][prow] = py; }
                __syncthreads();
                if (half == 0) {
                    out[(size_t)p * 3 + 0] = pv + s_part[0][prow] + th.bo;
                    out[(size_t)p * 3 + 1] = px + s_part[1][prow];
                    out[(size_t)p * 3 + 2] = py + s_part[2][prow];
                }
                break;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float* h = c == 0 ? hv : (c == 1 ? hx : hy);
                uint32_t hi[HP / 2], lo[HP / 2];
#pragma unroll
                for (int j = 0; j < HP / 2; ++j) { hi[j] = __float_as_uint(h[j]) & 0xffffe000u; lo[j] = __float_as_uint(h[j] - __uint_as_float(hi[j])); }
                const uint32_t a_hi = tb + lane_base + colAhi + c * KP + u0, a_lo = tb + lane_base + colAlo + c * KP + u0;
                uint32_t t8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) t8[j] = hi[j];
                hpv_tmem_st8(a_hi, t8); hpv_tmem_st2(a_hi + 8, hi[8], hi[9]);
#pragma unroll
                for (int j = 0; j < 8; ++j) t8[j] = lo[j];
                hpv_tmem_st8(a_lo, t8); hpv_tmem_st2(a_lo + 8, lo[8], lo[9]);
            }
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                hpv_tc_fence_after();
                if (hpv_elect_one()) {
                    const uint32_t bhi = hpv_smem_u32(&sB[0][0][0][0][0]) + l * (2 * (KP / 4) * NPAD * 16), blo = bhi + (KP / 4) * NPAD * 16;
                    if (tb == 0u) mlp3_issue<0, ND>(bhi, blo, bar); else mlp3_issue<256, ND>(bhi, blo, bar);
                }
                __syncwarp();
            }
        }
    }
    if (tid == 0 && cyc_out) cyc_out[blockIdx.x] = clock64() - t_start;
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 256);
}

// Issue-rate micro-benchmark: warp 0 issues `n` MMAs (M=128, N, K=8, TS) per round in a pattern over nD accumulators,
// commits, waits; cycles per MMA.  uniform != 0: the whole warp runs the loop, one elected lane issues.
__global__ void __launch_bounds__(128, 1) k_rate(int N, int nD, int n, int uniform, long long* out) {
    __shared__ __align__(128) uint32_t sB[3][2][32][4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) hpv_tmem_alloc(&tbase, 512);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < 3 * 2 * 32 * 4; i += 128) (&sB[0][0][0][0])[i] = 0u;
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    for (int c = 0; c < 512; c += 8) { uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; hpv_tmem_st8(tb + ((uint32_t)(warp * 32) << 16) + c, z); }
    hpv_tmem_wait_st();
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        hpv_tc_fence_after();
        const uint32_t idesc = hpv_umma_idesc_tf32(128, N, 0, 0);
        uint64_t bd[3];
        for (int i = 0; i < 3; ++i) bd[i] = hpv_umma_desc(hpv_smem_u32(&sB[i][0][0][0]), N * 16, 128);
        uint32_t phase = 0;
        long long best = 1ll << 60;
        for (int rep = 0; rep < 6; ++rep) {
            const long long t0 = clock64();
            if (uniform) {
                for (int i = 0; i < n; ++i) {
                    const uint32_t d = tb + 64 + (uint32_t)(i % nD) * 32, a = tb + (uint32_t)(i % 3) * 8;
                    if (hpv_elect_one()) hpv_umma_ts(d, a, bd[i % 3], idesc, 1);
                }
                if (hpv_elect_one()) hpv_umma_commit(&bar);
            } else if (tid == 0) {
                for (int i = 0; i < n; ++i) hpv_umma_ts(tb + 64 + (uint32_t)(i % nD) * 32, tb + (uint32_t)(i % 3) * 8, bd[i % 3], idesc, 1);
                hpv_umma_commit(&bar);
            }
            __syncwarp();
            const long long t1 = clock64();
            hpv_mbar_wait(&bar, phase);
            phase ^= 1;
            const long long t2 = clock64();
            if (t2 - t0 < best) { best = t2 - t0; if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; } }
        }
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, 512);
}


// Issue-rate micro-benchmark, second form: the CTA owns all 512 columns (base 0), every operand is a constant;
// one elected thread issues n = 36 MMAs per round over nD accumulators.
template <int ND, int NDACC>
__global__ void __launch_bounds__(128, 1) k_rate2(long long* out) {
    __shared__ __align__(128) uint32_t sB[3][2][64][4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) hpv_tmem_alloc(&tbase, 512);
    if (tid == 0) { hpv_mbar_init(&bar, 1); hpv_mbar_init_fence(); }
    for (int i = tid; i < 3 * 2 * 64 * 4; i += 128) (&sB[0][0][0][0])[i] = 0u;
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    if (tbase != 0u) { asm volatile("trap;"); }
    for (int c = 0; c < 512; c += 8) { uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0}; hpv_tmem_st8(((uint32_t)(warp * 32) << 16) + c, z); }
    hpv_tmem_wait_st();
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        hpv_tc_fence_after();
        constexpr uint32_t idesc = hpv_umma_idesc_tf32(128, ND, 0, 0);
        uint32_t phase = 0;
        long long best = 1ll << 60, best_issue = 0;
        for (int rep = 0; rep < 8; ++rep) {
            const long long t0 = clock64();
            if (hpv_elect_one()) {
#pragma unroll
                for (int i = 0; i < 36; ++i) {
                    const uint64_t bd = hpv_umma_desc(hpv_smem_u32(&sB[i % 3][0][0][0]), ND * 16, 128);
                    hpv_umma_ts(64u + (uint32_t)(i % NDACC) * 64u, (uint32_t)(i % 3) * 8u, bd, idesc, 1);
                }
                hpv_umma_commit(&bar);
            }
            __syncwarp();
            const long long t1 = clock64();
            hpv_mbar_wait(&bar, phase);
            phase ^= 1;
            const long long t2 = clock64();
            if (t2 - t0 < best) { best = t2 - t0; best_issue = t1 - t0; }
        }
        if (tid == 0) { out[0] = best_issue; out[1] = best; }
    }
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(0u, 512);
}

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
static float tf32_round(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; }

static int test_gemm(int mode, int N, int K, int split, bool pre_round) {
    std::vector<float> A(128 * K), B(N * K), D(128 * N, -777.0f);
    for (auto& v : A) v = (float)frand();
    for (auto& v : B) v = (float)frand();
    if (pre_round) { for (auto& v : A) v = tf32_round(v); for (auto& v : B) v = tf32_round(v); }
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(2 * 128 * K + 2 * N * K) * 4;
    CK(cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm<<<1, 128, smem>>>(dA, dB, dD, N, K, mode, split);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double emax = 0, rmax = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * (double)B[n * K + k];
            emax = fmax(emax, fabs(r - D[m * N + n])); rmax = fmax(rmax, fabs(r));
        }
    printf("gemm mode=%s N=%d K=%d split=%d pre_round=%d: max|err| %.3e of max|ref| %.3e (rel %.3e)  D[0][0]=%g D[127][N-1]=%g\n", mode ? "TS" : "SS", N, K,
           split, (int)pre_round, emax, rmax, emax / rmax, D[0], D[127 * N + N - 1]);
    return 0;
}

static int test_mn(int N) {
    const int K = 128, M = 64;
    std::vector<float> A(K * M), B(K * N), D(M * N, -777.0f);
    for (auto& v : A) v = tf32_round((float)frand());
    for (auto& v : B) v = tf32_round((float)frand());
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(M * K + N * K) * 4;
    CK(cudaFuncSetAttribute(k_gemm_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_mn<<<1, 128, smem>>>(dA, dB, dD, N, K);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double emax = 0, rmax = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[k * M + m] * (double)B[k * N + n];
            emax = fmax(emax, fabs(r - D[m * N + n])); rmax = fmax(rmax, fabs(r));
        }
    printf("gemm MN-major M=64 N=%d K=128: max|err| %.3e of max|ref| %.3e (rel %.3e)\n", N, emax, rmax, emax / rmax);
    return 0;
}

static void mlp_ref(const MlpTheta& t, double x, double y, double f[3]) {
    double hv[HP], hx[HP], hy[HP];
    for (int u = 0; u < HP; ++u) {
        const double z = t.b1[u] + x * t.W1[0][u] + y * t.W1[1][u], a = tanh(z), s1 = 1 - a * a;
        hv[u] = a; hx[u] = s1 * t.W1[0][u]; hy[u] = s1 * t.W1[1][u];
    }
    for (int l = 0; l < NHH; ++l) {
        double zv[HP], zx[HP], zy[HP];
        for (int n = 0; n < HP; ++n) {
            zv[n] = t.bh[l][n]; zx[n] = 0; zy[n] = 0;
            for (int k = 0; k < HP; ++k) { zv[n] += hv[k] * t.Wh[l][k][n]; zx[n] += hx[k] * t.Wh[l][k][n]; zy[n] += hy[k] * t.Wh[l][k][n]; }
        }
        for (int n = 0; n < HP; ++n) { const double a = tanh(zv[n]), s1 = 1 - a * a; hv[n] = a; hx[n] = s1 * zx[n]; hy[n] = s1 * zy[n]; }
    }
    f[0] = t.bo; f[1] = 0; f[2] = 0;
    for (int u = 0; u < HP; ++u) { f[0] += hv[u] * t.Wo[u]; f[1] += hx[u] * t.Wo[u]; f[2] += hy[u] * t.Wo[u]; }
}


static void match_rows(const std::vector<float>& Dall, const std::vector<double>& ref, int M, int N, const char* tag) {
    // for every reference row find the lane of the dump that matches it best
    int found = 0; double worst = 0; int lane_of[64];
    for (int m = 0; m < M; ++m) {
        double best = 1e30; int bl = -1;
        for (int l = 0; l < 128; ++l) {
            double e = 0, r = 0;
            for (int n = 0; n < N; ++n) { const double d = Dall[l * N + n]; e = fmax(e, std::isnan(d) ? 1e30 : fabs(d - ref[m * N + n])); r = fmax(r, fabs(ref[m * N + n])); }
            if (e / r < best) { best = e / r; bl = l; }
        }
        lane_of[m] = bl; if (best < 1e-3) ++found; worst = fmax(worst, best);
    }
    printf("%s: %d of %d rows found (worst rel err %.2e); lanes of rows 0,1,15,16,17,31,32,48,63: %d %d %d %d %d %d %d %d %d\n", tag, found, M, worst,
           lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[17], lane_of[31], lane_of[32], lane_of[48], lane_of[63]);
}

static int test_m64() {
    const int N = 32, K = 24, M = 64;
    std::vector<float> A(M * K), B(N * K), D(128 * N, 0.f);
    for (auto& v : A) v = tf32_round((float)frand());
    for (auto& v : B) v = tf32_round((float)frand());
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = (size_t)(M * K + N * K) * 4;
    k_gemm_m64<<<1, 128, smem>>>(dA, dB, dD, N, K);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    std::vector<double> ref(M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k]; ref[m * N + n] = r; }
    match_rows(D, ref, M, N, "M=64 K-major SS");
    return 0;
}

static int test_mn2() {
    const int K = 128, M = 64;
    for (int N : {32, 24}) {
        std::vector<float> A(K * M), B(K * N), D(128 * N, 0.f);
        for (auto& v : A) v = tf32_round((float)frand());
        for (auto& v : B) v = tf32_round((float)frand());
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        std::vector<double> ref(M * N);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[k * M + m] * B[k * N + n]; ref[m * N + n] = r; }
        const size_t smem = (size_t)(M * K + N * K) * 4;
        CK(cudaFuncSetAttribute(k_gemm_mn2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int lay = 0; lay < 2; ++lay) {
            const int sR4 = lay == 0 ? (K / 8) * 128 : 128, sK8 = lay == 0 ? 128 : (M / 4) * 128;
            for (int sw = 0; sw < 2; ++sw) {
                const int lbo = sw == 0 ? sK8 : sR4, sbo = sw == 0 ? sR4 : sK8;
                k_gemm_mn2<<<1, 128, smem>>>(dA, dB, dD, N, K, sR4, sK8, lbo, sbo, 0);
                CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
                CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
                char tag[160];
                snprintf(tag, sizeof(tag), "MN-major N=%d layout %s desc(lbo=%s, sbo=%s)", N, lay == 0 ? "k-blocks adjacent" : "row-blocks adjacent",
                         sw == 0 ? "k-block stride" : "row-block stride", sw == 0 ? "row-block stride" : "k-block stride");
                match_rows(D, ref, M, N, tag);
            }
        }
    }
    return 0;
}


static int test_mn3() {
    const int K = 128, M = 64;
    for (int N : {32, 24}) {
        std::vector<float> A(K * M), B(K * N), D(128 * N, 0.f);
        for (auto& v : A) v = tf32_round((float)frand());
        for (auto& v : B) v = tf32_round((float)frand());
        float *dA, *dB, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        std::vector<double> ref(M * N);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double r = 0; for (int k = 0; k < K; ++k) r += (double)A[k * M + m] * B[k * N + n]; ref[m * N + n] = r; }
        const size_t smem = (size_t)(K / 4) * (2 * 512 + 512) + 2048;
        CK(cudaFuncSetAttribute(k_gemm_mn3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        for (int variant = 0; variant < 2; ++variant) {
            k_gemm_mn3<<<1, 128, smem>>>(dA, dB, dD, N, K, variant);
            CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            char tag[160];
            snprintf(tag, sizeof(tag), "MN-major SW128/32B-base N=%d desc(%s)", N, variant == 0 ? "lbo=row-block stride, sbo=k-block stride" : "swapped");
            match_rows(D, ref, M, N, tag);
            printf("   D lane0: %g %g %g %g | ref row0: %g %g %g %g\n", D[0], D[1], D[2], D[3], ref[0], ref[1], ref[2], ref[3]);
        }
    }
    return 0;
}

static int test_rate() {
    long long* dout; CK(cudaMalloc(&dout, 16));
    for (int N : {32, 24, 64}) for (int uniform : {0, 1}) for (int nD : {1, 2, 3, 6, 9}) {
        const int n = 36;
        if ((nD * 32 + 64) > 512) continue;
        k_rate<<<1, 128>>>(N, nD, n, uniform, dout);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        long long h[2]; CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
        printf("rate N=%d %s nD=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%d MMAs: %lld / %lld cycles)\n", N, uniform ? "warp-uniform" : "thread-0", nD,
               (double)h[0] / n, (double)h[1] / n, n, h[0], h[1]);
    }
    return 0;
}

static int test_mlp() {
    MlpTheta t;
    float* tf = reinterpret_cast<float*>(&t);
    for (size_t i = 0; i < sizeof(MlpTheta) / 4; ++i) tf[i] = (float)(frand() * 0.45);
    const int n_tiles = 3200 * 8;                  // 8 x C3's 409 600 points
    MlpTheta* dth; float* dout; long long* dcyc;
    CK(cudaMalloc(&dth, sizeof(MlpTheta))); CK(cudaMalloc(&dout, (size_t)n_tiles * 128 * 3 * 4)); CK(cudaMalloc(&dcyc, 296 * 8));
    CK(cudaMemcpy(dth, &t, sizeof(MlpTheta), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int split = 1; split >= 0; --split) {
        for (int grid : {296, 148}) {
            CK(cudaMemset(dout, 0, (size_t)n_tiles * 128 * 3 * 4));
            k_mlp<<<grid, 256>>>(dth, dout, n_tiles, split, dcyc);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int r = 0; r < 5; ++r) k_mlp<<<grid, 256>>>(dth, dout, n_tiles, split, dcyc);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= 5;
            std::vector<float> out((size_t)n_tiles * 128 * 3);
            CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<long long> cyc(grid);
            CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
            double emax[3] = {0, 0, 0}, rmax[3] = {0, 0, 0};
            for (int p = 0; p < n_tiles * 128; p += 37) {
                const double x = -1.0 + 2.0 * (double)(float)(p % 1024) / 1024.0, y = -1.0 + 2.0 * (double)(float)(p / 1024 % 1024) / 1024.0;
                double f[3];
                mlp_ref(t, x, y, f);
                for (int c = 0; c < 3; ++c) { emax[c] = fmax(emax[c], fabs(f[c] - out[(size_t)p * 3 + c])); rmax[c] = fmax(rmax[c], fabs(f[c])); }
            }
            const double tiles_per_cta = (double)n_tiles / grid;
            printf("mlp split=%d grid=%d: %.1f us for %d points = %.4f ns/point (%.1f us per 409600 points) | %.0f cycles per tile (CTA), "
                   "%.0f per tile per SM | err u %.2e/%.2e ux %.2e/%.2e uy %.2e/%.2e\n",
                   split, grid, ms * 1e3, n_tiles * 128, ms * 1e6 / (n_tiles * 128.0), ms * 1e3 / 8, (double)cyc[0] / tiles_per_cta,
                   (double)cyc[0] / tiles_per_cta / (grid / 148), emax[0], rmax[0], emax[1], rmax[1], emax[2], rmax[2]);
        }
    }
    return 0;
}


template <int ORDER, int ND>
static int run_mlp2(const MlpTheta& t, MlpTheta* dth, float* dout, long long* dcyc, int n_tiles) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid : {296, 148}) {
        CK(cudaMemset(dout, 0, (size_t)n_tiles * 128 * 3 * 4));
        k_mlp2<ORDER, ND><<<grid, 256>>>(dth, dout, n_tiles, dcyc);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) k_mlp2<ORDER, ND><<<grid, 256>>>(dth, dout, n_tiles, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        std::vector<float> out((size_t)n_tiles * 128 * 3);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
        double emax[3] = {0, 0, 0}, rmax[3] = {0, 0, 0};
        for (int p = 0; p < n_tiles * 128; p += 37) {
            const double x = -1.0 + 2.0 * (double)(float)(p % 1024) / 1024.0, y = -1.0 + 2.0 * (double)(float)(p / 1024 % 1024) / 1024.0;
            double f[3]; mlp_ref(t, x, y, f);
            for (int c = 0; c < 3; ++c) { emax[c] = fmax(emax[c], fabs(f[c] - out[(size_t)p * 3 + c])); rmax[c] = fmax(rmax[c], fabs(f[c])); }
        }
        const double tiles_per_cta = (double)n_tiles / grid;
        printf("mlp2 order=%d N=%d grid=%d: %.1f us = %.4f ns/point (%.1f us per 409600 points) | %.0f cycles per tile (CTA), %.0f per tile per SM | "
               "err u %.2e/%.2e ux %.2e/%.2e uy %.2e/%.2e\n", ORDER, ND, grid, ms * 1e3, ms * 1e6 / (n_tiles * 128.0), ms * 1e3 / 8,
               (double)cyc[0] / tiles_per_cta, (double)cyc[0] / tiles_per_cta / (grid / 148), emax[0], rmax[0], emax[1], rmax[1], emax[2], rmax[2]);
    }
    return 0;
}

static int test_mlp2() {
    MlpTheta t;
    float* tf = reinterpret_cast<float*>(&t);
    for (size_t i = 0; i < sizeof(MlpTheta) / 4; ++i) tf[i] = (float)(frand() * 0.45);
    const int n_tiles = 3200 * 8;
    MlpTheta* dth; float* dout; long long* dcyc;
    CK(cudaMalloc(&dth, sizeof(MlpTheta))); CK(cudaMalloc(&dout, (size_t)n_tiles * 128 * 3 * 4)); CK(cudaMalloc(&dcyc, 296 * 8));
    CK(cudaMemcpy(dth, &t, sizeof(MlpTheta), cudaMemcpyHostToDevice));
    int r = run_mlp2<0, 32>(t, dth, dout, dcyc, n_tiles); if (r) return r;
    r = run_mlp2<1, 32>(t, dth, dout, dcyc, n_tiles); if (r) return r;
    r = run_mlp2<1, 24>(t, dth, dout, dcyc, n_tiles); if (r) return r;
    return 0;
}


template <int ND>
static int run_mlp3(const MlpTheta& t, MlpTheta* dth, float* dout, long long* dcyc, int n_tiles) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid : {296, 148}) {
        CK(cudaMemset(dout, 0, (size_t)n_tiles * 128 * 3 * 4));
        k_mlp3<ND><<<grid, 256>>>(dth, dout, n_tiles, dcyc);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 5; ++r) k_mlp3<ND><<<grid, 256>>>(dth, dout, n_tiles, dcyc);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        std::vector<float> out((size_t)n_tiles * 128 * 3);
        CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
        std::vector<long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
        double emax[3] = {0, 0, 0}, rmax[3] = {0, 0, 0};
        for (int p = 0; p < n_tiles * 128; p += 37) {
            const double x = -1.0 + 2.0 * (double)(float)(p % 1024) / 1024.0, y = -1.0 + 2.0 * (double)(float)(p / 1024 % 1024) / 1024.0;
            double f[3]; mlp_ref(t, x, y, f);
            for (int c = 0; c < 3; ++c) { emax[c] = fmax(emax[c], fabs(f[c] - out[(size_t)p * 3 + c])); rmax[c] = fmax(rmax[c], fabs(f[c])); }
        }
        const double tiles_per_cta = (double)n_tiles / grid;
        printf("mlp3 N=%d grid=%d: %.1f us = %.4f ns/point (%.1f us per 409600 points) | %.0f cycles per tile (CTA), %.0f per tile per SM | "
               "err u %.2e/%.2e ux %.2e/%.2e uy %.2e/%.2e\n", ND, grid, ms * 1e3, ms * 1e6 / (n_tiles * 128.0), ms * 1e3 / 8,
               (double)cyc[0] / tiles_per_cta, (double)cyc[0] / tiles_per_cta / (grid / 148), emax[0], rmax[0], emax[1], rmax[1], emax[2], rmax[2]);
    }
    return 0;
}

static int test_mlp3() {
    MlpTheta t;
    float* tf = reinterpret_cast<float*>(&t);
    for (size_t i = 0; i < sizeof(MlpTheta) / 4; ++i) tf[i] = (float)(frand() * 0.45);
    const int n_tiles = 3200 * 8;
    MlpTheta* dth; float* dout; long long* dcyc;
    CK(cudaMalloc(&dth, sizeof(MlpTheta))); CK(cudaMalloc(&dout, (size_t)n_tiles * 128 * 3 * 4)); CK(cudaMalloc(&dcyc, 296 * 8));
    CK(cudaMemcpy(dth, &t, sizeof(MlpTheta), cudaMemcpyHostToDevice));
    int r = run_mlp3<32>(t, dth, dout, dcyc, n_tiles); if (r) return r;
    return run_mlp3<24>(t, dth, dout, dcyc, n_tiles);
}

template <int ND, int NDACC>
static int run_rate2(long long* dout) {
    k_rate2<ND, NDACC><<<1, 128>>>(dout);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    long long h[2]; CK(cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost));
    printf("rate2 N=%d accumulators=%d: issue %.1f cyc/MMA, issue+complete %.1f cyc/MMA (36 MMAs: %lld / %lld cycles)\n", ND, NDACC, (double)h[0] / 36,
           (double)h[1] / 36, h[0], h[1]);
    return 0;
}
static int test_rate2() {
    long long* dout; CK(cudaMalloc(&dout, 16));
    int r = 0;
    r = r ? r : run_rate2<32, 1>(dout); r = r ? r : run_rate2<32, 3>(dout); r = r ? r : run_rate2<32, 6>(dout);
    r = r ? r : run_rate2<24, 1>(dout); r = r ? r : run_rate2<24, 3>(dout);
    r = r ? r : run_rate2<64, 1>(dout); r = r ? r : run_rate2<64, 3>(dout);
    r = r ? r : run_rate2<16, 3>(dout);
    return r;
}

int main(int argc, char** argv) {
    const int t = argc > 1 ? atoi(argv[1]) : 1;
    srand(1234);
    if (t == 1) return test_gemm(0, 32, 24, 0, true);
    if (t == 2) return test_gemm(1, 32, 24, 0, true);
    if (t == 3) return test_gemm(0, 24, 24, 0, true);
    if (t == 4) { int r = test_mn(32); return r ? r : test_mn(24); }
    if (t == 5) { int r = test_gemm(1, 32, 24, 1, false); if (r) return r; r = test_gemm(1, 32, 24, 0, false); return r ? r : test_gemm(0, 32, 24, 1, false); }
    if (t == 6) return test_mlp();
    if (t == 7) return test_gemm(1, 24, 24, 1, false);
    if (t == 8) return test_m64();
    if (t == 9) return test_mn2();
    if (t == 10) return test_rate();
    if (t == 11) return test_mlp2();
    if (t == 12) return test_mlp3();
    if (t == 14) return test_mn3();
    if (t == 13) return test_rate2();
    return 0;
}
