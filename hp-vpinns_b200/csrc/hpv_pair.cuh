// Packed FP32x2 arithmetic (sm_100 fma/mul/add.rn.f32x2 -> SASS FFMA2 / FMUL2 / FADD2): two lanes of work per
// issue slot.  The host emulation build uses a plain struct with the same per-lane rounding.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HPV_HD __host__ __device__ __forceinline__
#else
#define HPV_HD inline
#endif

#if defined(__CUDA_ARCH__)
typedef unsigned long long hpv_pair;
HPV_HD hpv_pair hpv_pack(float lo, float hi) {
    hpv_pair p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
}
HPV_HD void hpv_unpack(hpv_pair p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
HPV_HD void hpv_fma2(hpv_pair& c, hpv_pair a, hpv_pair b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(a), "l"(b)); }
HPV_HD hpv_pair hpv_fma2r(hpv_pair a, hpv_pair b, hpv_pair c) {
    hpv_pair r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
HPV_HD hpv_pair hpv_mul2(hpv_pair a, hpv_pair b) { hpv_pair r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HPV_HD hpv_pair hpv_add2(hpv_pair a, hpv_pair b) { hpv_pair r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
HPV_HD float hpv_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
HPV_HD float hpv_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// Two consecutive floats at an 8-byte aligned address as one 64-bit load (constant memory: one LDCU.64).
HPV_HD hpv_pair hpv_ld_pair(const float* p) { return *reinterpret_cast<const unsigned long long*>(p); }
#else
struct hpv_pair { float lo, hi; };
HPV_HD hpv_pair hpv_pack(float lo, float hi) { hpv_pair p; p.lo = lo; p.hi = hi; return p; }
HPV_HD void hpv_unpack(hpv_pair p, float& lo, float& hi) { lo = p.lo; hi = p.hi; }
HPV_HD void hpv_fma2(hpv_pair& c, hpv_pair a, hpv_pair b) { c.lo = fmaf(a.lo, b.lo, c.lo); c.hi = fmaf(a.hi, b.hi, c.hi); }
HPV_HD hpv_pair hpv_fma2r(hpv_pair a, hpv_pair b, hpv_pair c) { hpv_pair r; r.lo = fmaf(a.lo, b.lo, c.lo); r.hi = fmaf(a.hi, b.hi, c.hi); return r; }
HPV_HD hpv_pair hpv_mul2(hpv_pair a, hpv_pair b) { hpv_pair r; r.lo = a.lo * b.lo; r.hi = a.hi * b.hi; return r; }
HPV_HD hpv_pair hpv_add2(hpv_pair a, hpv_pair b) { hpv_pair r; r.lo = a.lo + b.lo; r.hi = a.hi + b.hi; return r; }
HPV_HD hpv_pair hpv_ld_pair(const float* p) { hpv_pair r; r.lo = p[0]; r.hi = p[1]; return r; }
HPV_HD float hpv_ex2(float x) { return exp2f(x); }
HPV_HD float hpv_rcp(float x) { return 1.0f / x; }
#endif

HPV_HD hpv_pair hpv_dup(float x) { return hpv_pack(x, x); }
