// Development probe (not part of the product): throughput of one 20x20 layer product of the MLP, two channels, in
// the two formulations discussed in DESIGN.md section 4:
//   A  thread per point, inputs from the thread's slot row, weights from constant memory through the uniform
//      datapath (LDCU.64 -> FFMA2 with a UR operand)  -- what hpv_matmul_slot does today;
//   B  warp-cooperative: the warp's 32 points x 20 units as a register tile of 4 points x 5 output units per lane,
//      activations stored [channel][unit][32 points], weights read from shared memory as two broadcast 128-bit
//      loads per input unit (pre-arranged per output-unit group), FFMA2 = scalar weight x pair of points.
// Each kernel applies NL layer products in sequence (output of one is the input of the next, scaled to stay finite)
// on n points; the time per point and layer product is printed.   nvcc -O3 -arch=sm_100a layer_product.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define HP 20
#define NCH 2
#define NL 8
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t pk(float a, float b) { pair_t p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(a), "f"(b)); return p; }
__device__ __forceinline__ void fma2(pair_t& c, pair_t a, pair_t b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }
__device__ __forceinline__ pair_t mul2(pair_t a, pair_t b) { pair_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

__constant__ __align__(16) float cW[NL * HP * HP];

// ---- A: thread per point ------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kernA(const float* __restrict__ in, float* __restrict__ out, int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slot = sm + warp * (NCH * 32 * HP);                  // [ch][32][HP]
    // CTA-uniform trip count (as in hpv_mlpbwd_body: a per-warp trip count would cost the uniform datapath)
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

// ---- A2: thread per TWO points (every LDCU.64 feeds 2 x NCH FFMA2) ---------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kernA2(const float* __restrict__ in, float* __restrict__ out, int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* slot = sm + warp * (2 * NCH * 32 * HP);              // [point 0/1][ch][32][HP]
    const int n_pairs = n_tiles / 2, n_grp = (n_pairs + WARPS - 1) / WARPS;
    for (int grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        const int pr_raw = grp * WARPS + warp;
        const bool valid = pr_raw < n_pairs;
        const int pr = valid ? pr_raw : n_pairs - 1;
        for (int q = 0; q < 2; ++q) {
            const int p = (2 * pr + q) * 32 + lane;
            for (int c = 0; c < NCH; ++c)
                for (int j = 0; j < HP; ++j) slot[((q * NCH + c) * 32 + lane) * HP + j] = in[((size_t)c * n_tiles * 32 + p) * HP + j];
        }
        pair_t acc[2][NCH][HP / 2];
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int m = 0; m < HP / 2; ++m) acc[q][c][m] = pk(0.f, 0.f);
            const float* W = cW + l * HP * HP;
            const float* row = slot + lane * HP;
#pragma unroll 1
            for (int i4 = 0; i4 < HP / 4; ++i4, row += 4, W += 4 * HP) {
                float4 x[2][NCH];
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int c = 0; c < NCH; ++c) x[q][c] = *reinterpret_cast<const float4*>(row + (q * NCH + c) * 32 * HP);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int m = 0; m < HP / 2; ++m) {
                        const pair_t w = *reinterpret_cast<const pair_t*>(W + k * HP + 2 * m);
#pragma unroll
                        for (int q = 0; q < 2; ++q)
#pragma unroll
                            for (int c = 0; c < NCH; ++c) {
                                const float xs = k == 0 ? x[q][c].x : (k == 1 ? x[q][c].y : (k == 2 ? x[q][c].z : x[q][c].w));
                                fma2(acc[q][c][m], pk(xs, xs), w);
                            }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int c = 0; c < NCH; ++c)
#pragma unroll
                    for (int j4 = 0; j4 < HP / 4; ++j4) {
                        ulonglong2 v; v.x = mul2(acc[q][c][2 * j4], pk(0.25f, 0.25f)); v.y = mul2(acc[q][c][2 * j4 + 1], pk(0.25f, 0.25f));
                        *reinterpret_cast<ulonglong2*>(slot + ((q * NCH + c) * 32 + lane) * HP + 4 * j4) = v;
                    }
        }
        if (valid)
            for (int q = 0; q < 2; ++q) {
                const int p = (2 * pr + q) * 32 + lane;
                for (int c = 0; c < NCH; ++c)
                    for (int j = 0; j < HP; ++j) out[((size_t)c * n_tiles * 32 + p) * HP + j] = slot[((q * NCH + c) * 32 + lane) * HP + j];
            }
    }
}

// ---- B: warp-cooperative -------------------------------------------------------------------------------------
// shared weights: Wg[l][b][i][8] = W_l[i][5b .. 5b+4], 0, 0, 0
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) kernB(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ Wg, int n_tiles) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;                                             // [NL][4][HP][8]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < NL * 4 * HP * 8; i += blockDim.x) sW[i] = Wg[i];
    __syncthreads();
    float* bufA = sm + NL * 4 * HP * 8 + warp * (2 * NCH * HP * 32);   // [ch][unit][32]
    float* bufB = bufA + NCH * HP * 32;
    const int a = lane >> 2, b = lane & 3;                       // point group (4 points), output-unit group (5 units)
    const int n_grp = (n_tiles + WARPS - 1) / WARPS;
    for (int grp = blockIdx.x; grp < n_grp; grp += gridDim.x) {
        const int tile_raw = grp * WARPS + warp;
        const bool valid = tile_raw < n_tiles;
        const int tile = valid ? tile_raw : n_tiles - 1;
        const int p = tile * 32 + lane;
        for (int c = 0; c < NCH; ++c)
            for (int j = 0; j < HP; ++j) bufA[(c * HP + j) * 32 + lane] = in[((size_t)c * n_tiles * 32 + p) * HP + j];
        __syncwarp();
        float* src = bufA; float* dst = bufB;
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            pair_t acc[NCH][5][2];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int u = 0; u < 5; ++u) { acc[c][u][0] = pk(0.f, 0.f); acc[c][u][1] = pk(0.f, 0.f); }
            const float* w = sW + ((l * 4 + b) * HP) * 8;
            const float* x = src + 4 * a;
#pragma unroll 4
            for (int i = 0; i < HP; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(w + i * 8);
                const float w4 = w[i * 8 + 4];
                const float ws[5] = {w0.x, w0.y, w0.z, w0.w, w4};
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const ulonglong2 xv = *reinterpret_cast<const ulonglong2*>(x + (c * HP + i) * 32);
#pragma unroll
                    for (int u = 0; u < 5; ++u) { fma2(acc[c][u][0], pk(ws[u], ws[u]), xv.x); fma2(acc[c][u][1], pk(ws[u], ws[u]), xv.y); }
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    ulonglong2 v; v.x = mul2(acc[c][u][0], pk(0.25f, 0.25f)); v.y = mul2(acc[c][u][1], pk(0.25f, 0.25f));
                    *reinterpret_cast<ulonglong2*>(dst + (c * HP + 5 * b + u) * 32 + 4 * a) = v;
                }
            __syncwarp();
            float* t = src; src = dst; dst = t;
        }
        if (valid)
            for (int c = 0; c < NCH; ++c)
                for (int j = 0; j < HP; ++j) out[((size_t)c * n_tiles * 32 + p) * HP + j] = src[(c * HP + j) * 32 + lane];
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
static void run(int n_tiles, const float* din, float* doutA, float* doutB, const float* dWg, const std::vector<float>& ref_check) {
    const int grid = 148;
    const size_t smA = (size_t)WARPS * NCH * 32 * HP * 4, smB = (size_t)(NL * 4 * HP * 8 + WARPS * 2 * NCH * HP * 32) * 4;
    cudaFuncSetAttribute(kernA<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smA);
    cudaFuncSetAttribute(kernB<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smB);
    const float ta = time_ms([&] { kernA<WARPS><<<grid, WARPS * 32, smA>>>(din, doutA, n_tiles); }, 10);
    const float tb = time_ms([&] { kernB<WARPS><<<grid, WARPS * 32, smB>>>(din, doutB, dWg, n_tiles); }, 10);
    cudaFuncSetAttribute(kernA2<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smA));
    const float ta2 = time_ms([&] { kernA2<WARPS><<<grid, WARPS * 32, 2 * smA>>>(din, doutB, n_tiles); }, 10);
    {
        const size_t n = (size_t)NCH * n_tiles * 32 * HP;
        std::vector<float> ha(n), hb(n);
        cudaMemcpy(ha.data(), doutA, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), doutB, n * 4, cudaMemcpyDeviceToHost);
        double md = 0;
        for (size_t i = 0; i < n; ++i) md = fmax(md, fabs((double)ha[i] - hb[i]));
        printf("warps/SM %2d  A2 (two points per thread)  %8.1f us = %.3f ns per point-layer | A2/A %.3f | max|A-A2| %.2e\n", WARPS, ta2 * 1e3, ta2 * 1e6 / ((double)n_tiles * 32 * NL), ta2 / ta, md);
    }
    time_ms([&] { kernB<WARPS><<<grid, WARPS * 32, smB>>>(din, doutB, dWg, n_tiles); }, 1);
    cudaError_t e = cudaDeviceSynchronize();
    const size_t n = (size_t)NCH * n_tiles * 32 * HP;
    std::vector<float> ha(n), hb(n);
    cudaMemcpy(ha.data(), doutA, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hb.data(), doutB, n * 4, cudaMemcpyDeviceToHost);
    double md = 0, mx = 0;
    for (size_t i = 0; i < n; ++i) { md = fmax(md, fabs((double)ha[i] - hb[i])); mx = fmax(mx, fabs((double)ha[i])); }
    const double prod = (double)n_tiles * 32 * NL;          // point-layer products (two channels each)
    printf("warps/SM %2d  A (thread per point, LDCU) %8.1f us = %.3f ns per point-layer | B (warp tile, LDS weights) %8.1f us = %.3f ns | B/A %.3f | max|A-B| %.2e of %.2e  [%s]\n",
           WARPS, ta * 1e3, ta * 1e6 / prod, tb * 1e3, tb * 1e6 / prod, tb / ta, md, mx, cudaGetErrorString(e));
    (void)ref_check;
}

int main() {
    const int n_tiles = 12800;                                // C3: 409 600 points
    std::vector<float> W(NL * HP * HP), Wg(NL * 4 * HP * 8, 0.f), x((size_t)NCH * n_tiles * 32 * HP);
    srand(1);
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : x) v = (rand() / (float)RAND_MAX - 0.5f);
    for (int l = 0; l < NL; ++l)
        for (int b = 0; b < 4; ++b)
            for (int i = 0; i < HP; ++i)
                for (int u = 0; u < 5; ++u) Wg[((l * 4 + b) * HP + i) * 8 + u] = W[l * HP * HP + i * HP + 5 * b + u];
    cudaMemcpyToSymbol(cW, W.data(), W.size() * 4);
    float *din, *da, *db, *dWg;
    cudaMalloc(&din, x.size() * 4); cudaMalloc(&da, x.size() * 4); cudaMalloc(&db, x.size() * 4); cudaMalloc(&dWg, Wg.size() * 4);
    cudaMemcpy(din, x.data(), x.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dWg, Wg.data(), Wg.size() * 4, cudaMemcpyHostToDevice);
    std::vector<float> dummy;
    run<8>(n_tiles, din, da, db, dWg, dummy);
    run<12>(n_tiles, din, da, db, dWg, dummy);
    run<16>(n_tiles, din, da, db, dWg, dummy);
    return 0;
}
