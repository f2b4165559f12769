// net_u and its input derivatives at scattered points (predict: P2D:255-257, P1D:197-199; boundary loss lossb:
// P2D:122, P1D:98, ADI:184; strong-form residual net_f / lossp: P2D:187-194, P1D:150-155, ADI:247-253).
// Thread per point, grid-stride.  When a target is given the kernel also forms the point residual
// r_i = sum_f c_f field_f(i) - target_i, its adjoint weight*2*r_i/n (input of the MLP reverse sweep) and the
// per-CTA partial sums of r_i^2.
#pragma once
#include "hpv_cta.cuh"
#include "hpv_const.cuh"
#include "hpv_slot.cuh"

template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_points_body(const HpvCta& c, const HpvPointArgs& a, float* gbar_out) {
    // shared memory: [reduction scratch: T floats][activation slot: NCH*T*SP floats]
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_red = sm;
    const int T = c.nthreads, tid = c.tid;
    float* s_slot = sm + T;
    const float* th = HPV_THETA(a.theta_pad);
    const float eps = a.eps[0];
    float cf[HPV_NFIELDS];
    for (int k = 0; k < HPV_NFIELDS; ++k) cf[k] = fmaf(eps, a.a1[k], a.a0[k]);
    hpv_sync(c);
    float sq = 0.0f;
    for (int base = c.bid * T; base < a.n; base += c.nblocks * T) {
        const int i = base + tid;
        if (i < a.n) {
            const float x = a.pts[(size_t)i * DIM];
            const float y = (DIM == 2) ? a.pts[(size_t)i * DIM + 1] : 0.0f;
            float f[HPV_NFIELDS];
            hpv_net_point_slot<DIM, MX, MY, HP, ACT>(th, a.nhid, a.off_wo, x, y,
                                                     s_slot + (tid >> 5) * (HpvMode<DIM, MX, MY>::NCH * 32 * HpvSP<HP>::value),
                                                     32, tid & 31, f);
            if (a.out_u) a.out_u[i] = f[0];
            if (a.out_d1) { a.out_d1[(size_t)i * DIM] = f[1]; if (DIM == 2) a.out_d1[(size_t)i * DIM + 1] = f[2]; }
            if (a.out_d2) { a.out_d2[(size_t)i * DIM] = f[3]; if (DIM == 2) a.out_d2[(size_t)i * DIM + 1] = f[4]; }
            if (a.target) {
                float r = -a.target[i];
#pragma unroll
                for (int k = 0; k < HPV_NFIELDS; ++k) r = fmaf(cf[k], f[k], r);
                if (a.resid) a.resid[i] = r;
                if (gbar_out) gbar_out[i] = a.weight * 2.0f * r / (float)a.n;
                sq = fmaf(r, r, sq);
            }
        }
    }
    if (a.target) {
        const float tot = hpv_block_sum(c, s_red, sq);
        if (tid == 0) a.blk_loss[c.bid] = a.weight * tot / (float)a.n;
    }
}
