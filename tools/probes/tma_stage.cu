// Development probe (not part of the product): staging the two test-function tables of a forward chunk
// (2 x 80 x 64 fp32 = 40 KB) from L2 into shared memory
//   (a) as the forward kernel does today: every thread copies 128-bit words through registers, four loads in flight;
//   (b) with the bulk-copy engine (TMA, 1-D): one elected thread issues cp.async.bulk ... mbarrier::complete_tx, the
//       CTA waits on the mbarrier.
// The shared region is first dirtied with ordinary stores (in the kernel it holds the activation slots before),
// so (b) includes the generic->async proxy fence it needs there.  Every wait has a clock-based timeout.
//   nvcc -O3 -arch=sm_100a tma_stage.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define TAB_FLOATS (80 * 64)
#define NTAB 2
#define THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(THREADS, 2) stage_kernel(const float* __restrict__ tabs, float* __restrict__ out, int iters, int use_tma, unsigned* err) {
    extern __shared__ __align__(128) unsigned char raw[];
    float* sm = reinterpret_cast<float*>(raw);                       // [NTAB][TAB_FLOATS]
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    unsigned phase = 0;
    for (int it = 0; it < iters; ++it) {
        // the region is in use by ordinary stores before each staging (activation slots in the real kernel)
        for (int i = tid; i < NTAB * TAB_FLOATS; i += THREADS) sm[i] = -1.0f;
        __syncthreads();
        if (use_tma) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes above -> async-proxy writes below
                const unsigned bytes = NTAB * TAB_FLOATS * 4;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(bytes) : "memory");
                for (int t = 0; t < NTAB; ++t)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(sm + t * TAB_FLOATS)), "l"(tabs + (size_t)t * TAB_FLOATS), "r"((unsigned)(TAB_FLOATS * 4)),
                                   "r"(smem_u32(&mbar)) : "memory");
            }
            unsigned done = 0;
            const long long t0 = clock64();
            while (!done) {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(smem_u32(&mbar)), "r"(phase) : "memory");
                if (!done && clock64() - t0 > 2000000000ll) { atomicExch(err, 1u); break; }
            }
            phase ^= 1;
        } else {
            const int n = TAB_FLOATS, step = THREADS * 4;
            for (int t = 0; t < NTAB; ++t) {
                const float* src = tabs + (size_t)t * TAB_FLOATS; float* dst = sm + t * TAB_FLOATS;
                int i = tid * 4;
                for (; i + 3 * step < n; i += 4 * step) {
                    const float4 v0 = *reinterpret_cast<const float4*>(src + i), v1 = *reinterpret_cast<const float4*>(src + i + step),
                                 v2 = *reinterpret_cast<const float4*>(src + i + 2 * step), v3 = *reinterpret_cast<const float4*>(src + i + 3 * step);
                    *reinterpret_cast<float4*>(dst + i) = v0; *reinterpret_cast<float4*>(dst + i + step) = v1;
                    *reinterpret_cast<float4*>(dst + i + 2 * step) = v2; *reinterpret_cast<float4*>(dst + i + 3 * step) = v3;
                }
                for (; i < n; i += step) *reinterpret_cast<float4*>(dst + i) = *reinterpret_cast<const float4*>(src + i);
            }
            __syncthreads();
        }
        // consume (a checksum over the staged tables, so that nothing is optimised away and the data is verified)
        for (int i = tid; i < NTAB * TAB_FLOATS; i += THREADS) acc += sm[i];
        __syncthreads();
    }
    out[blockIdx.x * THREADS + tid] = acc;
}

int main() {
    std::vector<float> h(NTAB * TAB_FLOATS);
    double want = 0;
    for (size_t i = 0; i < h.size(); ++i) { h[i] = (float)((i * 7919u) % 1013u) / 1013.0f; }
    float *d, *o; unsigned* derr;
    const int grid = 296, iters = 200;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, grid * THREADS * 4); cudaMalloc(&derr, 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice); cudaMemset(derr, 0, 4);
    const size_t smem = NTAB * TAB_FLOATS * 4;
    cudaFuncSetAttribute(stage_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> res[2];
    float ms[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        stage_kernel<<<grid, THREADS, smem>>>(d, o, 5, mode, derr);
        cudaEventRecord(e0);
        stage_kernel<<<grid, THREADS, smem>>>(d, o, iters, mode, derr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms[mode], e0, e1);
        res[mode].resize(grid * THREADS);
        cudaMemcpy(res[mode].data(), o, res[mode].size() * 4, cudaMemcpyDeviceToHost);
    }
    unsigned herr = 0; cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    double md = 0;
    for (size_t i = 0; i < res[0].size(); ++i) md = fmax(md, fabs((double)res[0][i] - res[1][i]));
    (void)want;
    printf("staging 2 x 20 KB tables per chunk, 296 CTAs x 256 threads, %d chunks (incl. dirtying + checksum of the region in both modes):\n", iters);
    printf("  threads (LDG.128 -> STS.128, 4 in flight) %.2f us per chunk\n  TMA bulk copy + mbarrier                    %.2f us per chunk\n", ms[0] * 1e3 / iters, ms[1] * 1e3 / iters);
    printf("  checksums equal: %s (max diff %.3g), timeout flag %u, last error: %s\n", md == 0 ? "yes" : "NO", md, herr, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
