// CTA execution context.  The kernel bodies are written against this small interface so that the SAME body
// runs as a CUDA kernel (threadIdx/blockIdx, __syncthreads, device atomics) and, for tests/emu only, on host
// threads with a barrier -- the thread emulation is a way to exercise the indexing/synchronisation logic of
// the kernels in the GPU-less build container; it is never a product path.
#pragma once
#include "hpv_math.cuh"

#if defined(__CUDA_ARCH__)
#define HPV_DEVICE_CODE 1
#else
#define HPV_DEVICE_CODE 0
#endif

struct HpvEmu;                                   // host emulation state (tests/emu/hpv_emu.h)
#if !HPV_DEVICE_CODE
void hpv_emu_barrier(HpvEmu* e);                 // provided by the emulation harness only
void hpv_emu_warp_barrier(HpvEmu* e, int warp);
float hpv_emu_shfl_xor(HpvEmu* e, int tid, float v, int mask);
#endif

struct HpvCta {
    int tid, nthreads, bid, nblocks;
    unsigned char* smem;
    HpvEmu* emu;
};

HPV_HD void hpv_sync(const HpvCta& c) {
#if HPV_DEVICE_CODE
    __syncthreads();
#else
    hpv_emu_barrier(c.emu);
#endif
}

// Programmatic dependent launch (the kernels of one training step are chained with
// cudaLaunchAttributeProgrammaticStreamSerialization, see hpv_launch_pdl): hpv_pdl_trigger lets the next kernel of
// the stream start its CTAs while this grid is still running; hpv_pdl_wait blocks until the previous kernel of the
// stream has completed and its writes are visible.  Every kernel calls the wait before it touches anything the
// previous kernel produced (no-ops when launched without the attribute, and in the host emulation).
HPV_HD void hpv_pdl_wait() {
#if HPV_DEVICE_CODE
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
HPV_HD void hpv_pdl_trigger() {
#if HPV_DEVICE_CODE
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// Warp-level synchronisation and exchange (every lane of the warp must call them).
HPV_HD void hpv_syncwarp(const HpvCta& c) {
#if HPV_DEVICE_CODE
    __syncwarp();
#else
    hpv_emu_warp_barrier(c.emu, c.tid >> 5);
#endif
}

HPV_HD float hpv_shfl_xor(const HpvCta& c, float v, int mask) {
#if HPV_DEVICE_CODE
    return __shfl_xor_sync(0xffffffffu, v, mask);
#else
    return hpv_emu_shfl_xor(c.emu, c.tid, v, mask);
#endif
}

HPV_HD unsigned int hpv_atomic_inc(unsigned int* p) {      // returns the previous value
#if HPV_DEVICE_CODE
    return atomicAdd(p, 1u);
#else
    return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL);
#endif
}

HPV_HD void hpv_fence() {
#if HPV_DEVICE_CODE
    __threadfence();
#else
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
#endif
}

// Loads of data written by OTHER CTAs of the same launch (L2, bypassing the non-coherent L1).
HPV_HD float hpv_ld_cg(const float* p) {
#if HPV_DEVICE_CODE
    return __ldcg(p);
#else
    return *reinterpret_cast<const volatile float*>(p);
#endif
}
HPV_HD HpvF4 hpv_ld4_cg(const float* p) {
#if HPV_DEVICE_CODE
    float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    HpvF4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
    return r;
#else
    HpvF4 r;
    const volatile float* q = p;
    r.x = q[0]; r.y = q[1]; r.z = q[2]; r.w = q[3];
    return r;
#endif
}

// Deterministic block-wide sum (fixed tree in shared memory); every thread gets the result.
// `red` must hold nthreads floats (any block size).
HPV_HD float hpv_block_sum(const HpvCta& c, float* red, float v) {
    hpv_sync(c);
    red[c.tid] = v;
    hpv_sync(c);
    int top = 1;
    while (top < c.nthreads) top <<= 1;
    for (int s = top >> 1; s > 0; s >>= 1) {
        if (c.tid < s && c.tid + s < c.nthreads) red[c.tid] += red[c.tid + s];
        hpv_sync(c);
    }
    float r = red[0];
    return r;
}

HPV_HD int hpv_align4(int n) { return (n + 3) & ~3; }
