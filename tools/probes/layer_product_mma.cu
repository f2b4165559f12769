// Development probe (not part of the product): the 20x20 layer product of the MLP, two channels, on the warp-level
// tensor-core path -- mma.sync.m16n8k8 TF32 with the 3-term split (A_lo.B_hi + A_hi.B_lo + A_hi.B_hi), fragments loaded
// from the warp's own activation rows in shared memory ([channel][32 points][20], the layout of the reverse sweep's
// slots) and from pre-split weights in shared memory -- against variant A of layer_product.cu (thread per point,
// weights from constant memory, FFMA2).  NL chained products on n points; time per point and layer product, and the
// difference of the results.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a layer_product_mma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define HP 20
#define NCH 2
#define NL 8
#define WS 24                                  // padded weight row stride / rows (K and N padded to 24)
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t pk(float a, float b) { pair_t p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b)); return p; }
__device__ __forceinline__ void fma2(pair_t& c, pair_t a, pair_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) { pair_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__constant__ __align__(16) float cW[NL * HP * HP];

// ---- A: thread per point (as layer_product.cu) ---------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kernA(const float* __restrict__ in, float* __restrict__ out, int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slot = sm + warp * (NCH * 32 * HP);
    const int n_grp = (n_tiles + WARPS - 1) / WARPS;
    for (int grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        const int tile_raw = grp * WARPS + warp;
        const bool valid = tile_raw < n_tiles;
        const int tile = valid ? tile_raw : n_tiles - 1;
        const int p = tile * 32 + lane;
        for (int c = 0; c < NCH; ++c)
            for (int j = 0; j < HP; ++j) slot[(c * 32 + lane) * HP + j] = in[((size_t)c * n_tiles * 32 + p) * HP + j];
        pair_t acc[NCH][HP / 2];
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int m = 0; m < HP / 2; ++m) acc[c][m] = pk(0.f, 0.f);
            const float* W = cW + l * HP * HP;
            const float* row = slot + lane * HP;
#pragma unroll 1
            for (int i4 = 0; i4 < HP / 4; ++i4, row += 4, W += 4 * HP) {
                float4 x[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) x[c] = *reinterpret_cast<const float4*>(row + c * 32 * HP);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int m = 0; m < HP / 2; ++m) {
                        const pair_t w = *reinterpret_cast<const pair_t*>(W + k * HP + 2 * m);
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            const float xs = k == 0 ? x[c].x : (k == 1 ? x[c].y : (k == 2 ? x[c].z : x[c].w));
                            fma2(acc[c][m], pk(xs, xs), w);
                        }
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int j4 = 0; j4 < HP / 4; ++j4) {
                    ulonglong2 v; v.x = mul2(acc[c][2 * j4], pk(0.25f, 0.25f)); v.y = mul2(acc[c][2 * j4 + 1], pk(0.25f, 0.25f));
                    *reinterpret_cast<ulonglong2*>(slot + (c * 32 + lane) * HP + 4 * j4) = v;
                }
        }
        if (valid)
            for (int c = 0; c < NCH; ++c)
                for (int j = 0; j < HP; ++j) out[((size_t)c * n_tiles * 32 + p) * HP + j] = slot[(c * 32 + lane) * HP + j];
    }
}

// ---- C: mma.sync m16n8k8 TF32, 3-term split -------------------------------------------------------------------------
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// Wsplit (global): [NL][2 (hi, lo)][WS k][WS n]
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kernC(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ Wsplit, int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;                                             // [NL][2][WS][WS]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < NL * 2 * WS * WS; i += blockDim.x) sW[i] = Wsplit[i];
    __syncthreads();
    float* slot = sm + NL * 2 * WS * WS + warp * (NCH * 32 * HP);
    const int n_grp = (n_tiles + WARPS - 1) / WARPS;
    for (int grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        const int tile_raw = grp * WARPS + warp;
        const bool valid = tile_raw < n_tiles;
        const int tile = valid ? tile_raw : n_tiles - 1;
        const int p = tile * 32 + lane;
        for (int c = 0; c < NCH; ++c)
            for (int j = 0; j < HP; ++j) slot[(c * 32 + lane) * HP + j] = in[((size_t)c * n_tiles * 32 + p) * HP + j];
        __syncwarp();
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            float acc[2][NCH][3][4];
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int n = 0; n < 3; ++n)
#pragma unroll
                        for (int q = 0; q < 4; ++q) acc[m][c][n][q] = 0.0f;
            const float* Whi = sW + (l * 2 + 0) * WS * WS;
            const float* Wlo = sW + (l * 2 + 1) * WS * WS;
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                unsigned ahi[2][NCH][4], alo[2][NCH][4];
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        const float* r0 = slot + (c * 32 + 16 * m + g) * HP + 8 * ks + t;
                        const float a0 = r0[0], a1 = r0[8 * HP];
                        const float a2 = (ks < 2) ? r0[4] : 0.0f, a3 = (ks < 2) ? r0[8 * HP + 4] : 0.0f;     // units 20..23 do not exist
                        split(a0, ahi[m][c][0], alo[m][c][0]); split(a1, ahi[m][c][1], alo[m][c][1]);
                        split(a2, ahi[m][c][2], alo[m][c][2]); split(a3, ahi[m][c][3], alo[m][c][3]);
                    }
#pragma unroll
                for (int n = 0; n < 3; ++n) {
                    const int o0 = (8 * ks + t) * WS + 8 * n + g, o1 = o0 + 4 * WS;
                    const unsigned bh0 = __float_as_uint(Whi[o0]), bh1 = __float_as_uint(Whi[o1]);
                    const unsigned bl0 = __float_as_uint(Wlo[o0]), bl1 = __float_as_uint(Wlo[o1]);
#pragma unroll
                    for (int m = 0; m < 2; ++m)
#pragma unroll
                        for (int c = 0; c < NCH; ++c) {
                            mma_tf32(acc[m][c][n], alo[m][c], bh0, bh1);
                            mma_tf32(acc[m][c][n], ahi[m][c], bl0, bl1);
                            mma_tf32(acc[m][c][n], ahi[m][c], bh0, bh1);
                        }
                }
            }
            __syncwarp();                                        // every lane has read its inputs of this layer
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int n = 0; n < 3; ++n) {
                        const int col = 8 * n + 2 * t;
                        if (col < HP) {
                            float* r0 = slot + (c * 32 + 16 * m + g) * HP + col;
                            *reinterpret_cast<float2*>(r0) = make_float2(0.25f * acc[m][c][n][0], 0.25f * acc[m][c][n][1]);
                            *reinterpret_cast<float2*>(r0 + 8 * HP) = make_float2(0.25f * acc[m][c][n][2], 0.25f * acc[m][c][n][3]);
                        }
                    }
            __syncwarp();
        }
        if (valid)
            for (int c = 0; c < NCH; ++c)
                for (int j = 0; j < HP; ++j) out[((size_t)c * n_tiles * 32 + p) * HP + j] = slot[(c * 32 + lane) * HP + j];
        __syncwarp();
    }
}

template <typename F>
static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <int WARPS>
static void run(int n_tiles, const float* din, float* doutA, float* doutC, const float* dWs, const std::vector<double>& ref, const std::vector<float>& x) {
    const int grid = 148;
    const size_t smA = (size_t)WARPS * NCH * 32 * HP * 4, smC = (size_t)(NL * 2 * WS * WS + WARPS * NCH * 32 * HP) * 4;
    cudaFuncSetAttribute(kernA<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    cudaFuncSetAttribute(kernC<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smC);
    const float ta = time_ms([&] { kernA<WARPS><<<grid, WARPS * 32, smA>>>(din, doutA, n_tiles); }, 10);
    const float tc = time_ms([&] { kernC<WARPS><<<grid, WARPS * 32, smC>>>(din, doutC, dWs, n_tiles); }, 10);
    cudaError_t e = cudaDeviceSynchronize();
    const size_t n = (size_t)NCH * n_tiles * 32 * HP;
    std::vector<float> ha(n), hc(n);
    cudaMemcpy(ha.data(), doutA, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hc.data(), doutC, n * 4, cudaMemcpyDeviceToHost);
    double md = 0, mx = 0, ea = 0, ec = 0;
    for (size_t i = 0; i < n; ++i) { md = fmax(md, fabs((double)ha[i] - hc[i])); mx = fmax(mx, fabs((double)ha[i])); }
    // float64 reference for the first 64 points of each channel
    for (int c = 0; c < NCH; ++c)
        for (int p = 0; p < 64; ++p)
            for (int j = 0; j < HP; ++j) {
                const size_t i = ((size_t)c * n_tiles * 32 + p) * HP + j;
                const double r = ref[((size_t)c * 64 + p) * HP + j];
                ea = fmax(ea, fabs(r - ha[i])); ec = fmax(ec, fabs(r - hc[i]));
            }
    const double prod = (double)n_tiles * 32 * NL;
    printf("warps/SM %2d  A (FFMA2, LDCU) %8.1f us = %.3f ns per point-layer | C (mma.sync 3xTF32) %8.1f us = %.3f ns | C/A %.3f | max|A-C| %.2e of %.2e | vs float64: A %.2e  C %.2e  [%s]\n",
           WARPS, ta * 1e3, ta * 1e6 / prod, tc * 1e3, tc * 1e6 / prod, tc / ta, md, mx, ea, ec, cudaGetErrorString(e));
    (void)x;
}

int main() {
    const int n_tiles = 12800;                                // C3: 409 600 points
    std::vector<float> W(NL * HP * HP), Ws((size_t)NL * 2 * WS * WS, 0.f), x((size_t)NCH * n_tiles * 32 * HP);
    srand(1);
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : x) v = (rand() / (float)RAND_MAX - 0.5f);
    for (int l = 0; l < NL; ++l)
        for (int k = 0; k < HP; ++k)
            for (int n = 0; n < HP; ++n) {
                const float w = W[l * HP * HP + k * HP + n];
                unsigned u; memcpy(&u, &w, 4); u &= 0xffffe000u;
                float hi; memcpy(&hi, &u, 4);
                Ws[((size_t)(l * 2 + 0) * WS + k) * WS + n] = hi;
                Ws[((size_t)(l * 2 + 1) * WS + k) * WS + n] = w - hi;
            }
    // float64 reference of the chain for the first 64 points
    std::vector<double> ref((size_t)NCH * 64 * HP);
    for (int c = 0; c < NCH; ++c)
        for (int p = 0; p < 64; ++p) {
            double h[HP], z[HP];
            for (int j = 0; j < HP; ++j) h[j] = x[((size_t)c * n_tiles * 32 + p) * HP + j];
            for (int l = 0; l < NL; ++l) {
                for (int n = 0; n < HP; ++n) { z[n] = 0; for (int k = 0; k < HP; ++k) z[n] += h[k] * (double)W[l * HP * HP + k * HP + n]; }
                for (int n = 0; n < HP; ++n) h[n] = 0.25 * z[n];
            }
            for (int j = 0; j < HP; ++j) ref[((size_t)c * 64 + p) * HP + j] = h[j];
        }
    cudaMemcpyToSymbol(cW, W.data(), W.size() * 4);
    float *din, *da, *dc, *dWs;
    cudaMalloc(&din, x.size() * 4); cudaMalloc(&da, x.size() * 4); cudaMalloc(&dc, x.size() * 4); cudaMalloc(&dWs, Ws.size() * 4);
    cudaMemcpy(din, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dWs, Ws.data(), Ws.size() * 4, cudaMemcpyHostToDevice);
    run<8>(n_tiles, din, da, dc, dWs, ref, x);
    run<12>(n_tiles, din, da, dc, dWs, ref, x);
    run<16>(n_tiles, din, da, dc, dWs, ref, x);
    return 0;
}
