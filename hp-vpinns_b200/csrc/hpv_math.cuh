// Per-point MLP arithmetic of the hp-VPINN hot path: net_u (P2D:158-173, P1D:128-142, ADI:219-234) and its
// input derivatives (tf.gradients chains of P2D:175-185, P1D:144-148, ADI:236-245) in analytic forward mode,
// plus the hand-derived reverse sweep through that forward-mode computation (what TF builds for
// AdamOptimizer.minimize, P2D:131-132).  fp32, FFMA.  Compiles for the device (nvcc) and, for the
// thread-emulated kernel tests under tests/emu, for the host (g++).
#pragma once
#include <math.h>
#include "hpv_types.h"

#if defined(__CUDACC__)
#define HPV_HD __host__ __device__ __forceinline__
#else
#define HPV_HD inline
#endif

enum { HPV_ACT_SIN = 0, HPV_ACT_TANH = 1 };

struct alignas(16) HpvF4 { float x, y, z, w; };

HPV_HD HpvF4 hpv_ld4(const float* p) { return *reinterpret_cast<const HpvF4*>(p); }
HPV_HD void hpv_st4(float* p, const HpvF4& v) { *reinterpret_cast<HpvF4*>(p) = v; }

// ---- padded parameter layout --------------------------------------------------------------------------
// [W1: DIM x HP][b1: HP] { [W_l: HP x HP][b_l: HP] } x (nhid-1) [Wo: HP][bo, 0, 0, 0]
// Hidden width H is zero-padded to HP: a padded unit has z = 0 -> sigma(0) = 0 for both sin and tanh and
// all its tangents are 0, so the padding is exact.
HPV_HD int hpv_off_w1() { return 0; }
HPV_HD int hpv_off_b1(int dim, int hp) { return dim * hp; }
HPV_HD int hpv_off_wl(int dim, int hp, int l /*1..nhid-1*/) { return (dim + 1) * hp + (l - 1) * (hp * hp + hp); }
HPV_HD int hpv_off_wo(int dim, int hp, int nhid) { return (dim + 1) * hp + (nhid - 1) * (hp * hp + hp); }
HPV_HD int hpv_theta_pad_n(int dim, int hp, int nhid) { return hpv_off_wo(dim, hp, nhid) + hp + 4; }

// ---- activations ----------------------------------------------------------------------------------------
// tanh as 1 - 2/(exp(2z)+1): absolute error ~1e-7 (what matters: the value feeds dot products), saturates
// cleanly to +-1, no branches.
HPV_HD float hpv_tanh(float z) {
#if defined(__CUDA_ARCH__)
    float e = exp2f(z * 2.8853900817779268f);          // exp(2z)
    return 1.0f - __fdividef(2.0f, e + 1.0f);
#else
    float e = expf(2.0f * z);
    return 1.0f - 2.0f / (e + 1.0f);
#endif
}

template <int ACT>
HPV_HD void hpv_act(float z, float& a, float& s1, float& s2) {
    if constexpr (ACT == HPV_ACT_TANH) {
        a = hpv_tanh(z);
        s1 = fmaf(-a, a, 1.0f);
        s2 = -2.0f * a * s1;
    } else {
        float s, c;
#if defined(__CUDA_ARCH__)
        sincosf(z, &s, &c);
#else
        s = sinf(z); c = cosf(z);
#endif
        a = s; s1 = c; s2 = -s;
    }
}

template <int ACT>
HPV_HD float hpv_act_s3(float a, float s1) {       // third derivative from value and first derivative
    if constexpr (ACT == HPV_ACT_TANH) return -2.0f * s1 * fmaf(-3.0f * a, a, 1.0f);
    return -s1;
}

// ---- derivative modes -------------------------------------------------------------------------------------
// MX, MY in {0: value only, 1: first derivative, 2: first and second} along x and y (t).
template <int DIM, int MX, int MY>
struct HpvMode {
    static constexpr bool DX = MX >= 1;
    static constexpr bool DY = (DIM == 2) && (MY >= 1);
    static constexpr bool EX = MX >= 2;
    static constexpr bool EY = (DIM == 2) && (MY >= 2);
    static constexpr int NCH = 1 + (DX ? 1 : 0) + (DY ? 1 : 0) + (EX ? 1 : 0) + (EY ? 1 : 0);
    // channel slots in stored tiles: v, dx, dy, ex, ey in this order, skipping the absent ones
    static constexpr int C_V = 0;
    static constexpr int C_DX = 1;
    static constexpr int C_DY = 1 + (DX ? 1 : 0);
    static constexpr int C_EX = C_DY + (DY ? 1 : 0);
    static constexpr int C_EY = C_EX + (EX ? 1 : 0);
};

template <int HP, bool ON>
struct HpvVec { float a[ON ? HP : 1]; };

// Activations of one layer at one point: value and the carried tangents.
template <int DIM, int MX, int MY, int HP>
struct HpvState {
    typedef HpvMode<DIM, MX, MY> M;
    HpvVec<HP, true> v;
    HpvVec<HP, M::DX> dx;
    HpvVec<HP, M::DY> dy;
    HpvVec<HP, M::EX> ex;
    HpvVec<HP, M::EY> ey;
};

// First layer, pre-activations: z = b1 + x W1[0,:] (+ y W1[1,:]); dz/dx = W1[0,:]; d2z = 0.
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_layer1_pre(const float* th, float x, float y, HpvState<DIM, MX, MY, HP>& z) {
    typedef HpvMode<DIM, MX, MY> M;
    const float* W1 = th + hpv_off_w1();
    const float* b1 = th + hpv_off_b1(DIM, HP);
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvF4 wx = hpv_ld4(W1 + 4 * j4), b = hpv_ld4(b1 + 4 * j4);
        HpvF4 wy = wx;
        if constexpr (DIM == 2) wy = hpv_ld4(W1 + HP + 4 * j4);
        float wxs[4] = {wx.x, wx.y, wx.z, wx.w}, wys[4] = {wy.x, wy.y, wy.z, wy.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int j = 4 * j4 + k;
            float zz = fmaf(x, wxs[k], bs[k]);
            if constexpr (DIM == 2) zz = fmaf(y, wys[k], zz);
            z.v.a[j] = zz;
            if constexpr (M::DX) z.dx.a[j] = wxs[k];
            if constexpr (M::DY) z.dy.a[j] = wys[k];
            if constexpr (M::EX) z.ex.a[j] = 0.0f;
            if constexpr (M::EY) z.ey.a[j] = 0.0f;
        }
    }
}

// Pre-activations (z, dz, d2z) -> post-activations (h, dh, d2h), in place.
//   h = s(z);  dh = s'(z) dz;  d2h = s''(z) dz^2 + s'(z) d2z
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_activate(HpvState<DIM, MX, MY, HP>& s) {
    typedef HpvMode<DIM, MX, MY> M;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        float a, s1, s2;
        hpv_act<ACT>(s.v.a[j], a, s1, s2);
        s.v.a[j] = a;
        if constexpr (M::EX) s.ex.a[j] = fmaf(s2 * s.dx.a[j], s.dx.a[j], s1 * s.ex.a[j]);
        if constexpr (M::EY) s.ey.a[j] = fmaf(s2 * s.dy.a[j], s.dy.a[j], s1 * s.ey.a[j]);
        if constexpr (M::DX) s.dx.a[j] = s1 * s.dx.a[j];
        if constexpr (M::DY) s.dy.a[j] = s1 * s.dy.a[j];
    }
}

// Hidden layer l >= 1: out = in . W + b for the value channel, out = in . W for every tangent channel.
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_matmul(const float* W, const float* b, const HpvState<DIM, MX, MY, HP>& in,
                       HpvState<DIM, MX, MY, HP>& out) {
    typedef HpvMode<DIM, MX, MY> M;
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvF4 bb = hpv_ld4(b + 4 * j4);
        out.v.a[4 * j4 + 0] = bb.x; out.v.a[4 * j4 + 1] = bb.y; out.v.a[4 * j4 + 2] = bb.z; out.v.a[4 * j4 + 3] = bb.w;
    }
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        if constexpr (M::DX) out.dx.a[j] = 0.0f;
        if constexpr (M::DY) out.dy.a[j] = 0.0f;
        if constexpr (M::EX) out.ex.a[j] = 0.0f;
        if constexpr (M::EY) out.ey.a[j] = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < HP; ++i) {
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) {
            HpvF4 w = hpv_ld4(W + i * HP + 4 * j4);
            float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = 4 * j4 + k;
                out.v.a[j] = fmaf(in.v.a[i], ws[k], out.v.a[j]);
                if constexpr (M::DX) out.dx.a[j] = fmaf(in.dx.a[i], ws[k], out.dx.a[j]);
                if constexpr (M::DY) out.dy.a[j] = fmaf(in.dy.a[i], ws[k], out.dy.a[j]);
                if constexpr (M::EX) out.ex.a[j] = fmaf(in.ex.a[i], ws[k], out.ex.a[j]);
                if constexpr (M::EY) out.ey.a[j] = fmaf(in.ey.a[i], ws[k], out.ey.a[j]);
            }
        }
    }
}

// Transposed product for the reverse sweep: out[i] = sum_j in[j] * W[i][j]  (all channels alike, no bias).
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_matmul_t(const float* W, const HpvState<DIM, MX, MY, HP>& in, HpvState<DIM, MX, MY, HP>& out) {
    typedef HpvMode<DIM, MX, MY> M;
#pragma unroll
    for (int i = 0; i < HP; ++i) {
        float av = 0.f, adx = 0.f, ady = 0.f, aex = 0.f, aey = 0.f;
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) {
            HpvF4 w = hpv_ld4(W + i * HP + 4 * j4);
            float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int j = 4 * j4 + k;
                av = fmaf(in.v.a[j], ws[k], av);
                if constexpr (M::DX) adx = fmaf(in.dx.a[j], ws[k], adx);
                if constexpr (M::DY) ady = fmaf(in.dy.a[j], ws[k], ady);
                if constexpr (M::EX) aex = fmaf(in.ex.a[j], ws[k], aex);
                if constexpr (M::EY) aey = fmaf(in.ey.a[j], ws[k], aey);
            }
        }
        out.v.a[i] = av;
        if constexpr (M::DX) out.dx.a[i] = adx;
        if constexpr (M::DY) out.dy.a[i] = ady;
        if constexpr (M::EX) out.ex.a[i] = aex;
        if constexpr (M::EY) out.ey.a[i] = aey;
    }
}

// Output layer (linear): fields (u, u_x, u_y, u_xx, u_yy); absent ones are 0.
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_output(const float* Wo, const HpvState<DIM, MX, MY, HP>& h, float f[HPV_NFIELDS]) {
    typedef HpvMode<DIM, MX, MY> M;
    float u = Wo[HP], ux = 0.f, uy = 0.f, uxx = 0.f, uyy = 0.f;
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvF4 w = hpv_ld4(Wo + 4 * j4);
        float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int j = 4 * j4 + k;
            u = fmaf(h.v.a[j], ws[k], u);
            if constexpr (M::DX) ux = fmaf(h.dx.a[j], ws[k], ux);
            if constexpr (M::DY) uy = fmaf(h.dy.a[j], ws[k], uy);
            if constexpr (M::EX) uxx = fmaf(h.ex.a[j], ws[k], uxx);
            if constexpr (M::EY) uyy = fmaf(h.ey.a[j], ws[k], uyy);
        }
    }
    f[0] = u; f[1] = ux; f[2] = uy; f[3] = uxx; f[4] = uyy;
}

// Whole network at one point.
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_net_point(const float* th, int nhid, float x, float y, float f[HPV_NFIELDS]) {
    HpvState<DIM, MX, MY, HP> a, b;
    hpv_layer1_pre<DIM, MX, MY, HP>(th, x, y, a);
    hpv_activate<DIM, MX, MY, HP, ACT>(a);
    for (int l = 1; l < nhid; ++l) {
        const float* W = th + hpv_off_wl(DIM, HP, l);
        hpv_matmul<DIM, MX, MY, HP>(W, W + HP * HP, a, b);
        hpv_activate<DIM, MX, MY, HP, ACT>(b);
        a = b;
    }
    hpv_output<DIM, MX, MY, HP>(th + hpv_off_wo(DIM, HP, nhid), a, f);
}

// Reverse of hpv_activate.  In: pre-activations z (value + tangents) and the adjoints of the post-activations
// (g).  Out (in place in g): adjoints of the pre-activations.
//   zbar   = hbar s1 + sum_d [ dhbar_d s2 dz_d + d2hbar_d (s3 dz_d^2 + s2 d2z_d) ]
//   dzbar  = dhbar s1 + 2 d2hbar s2 dz
//   d2zbar = d2hbar s1
// Also returns the post-activations in z (needed as the left factor of the weight gradient one layer up).
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_activate_bwd(HpvState<DIM, MX, MY, HP>& z, HpvState<DIM, MX, MY, HP>& g) {
    typedef HpvMode<DIM, MX, MY> M;
#pragma unroll
    for (int j = 0; j < HP; ++j) {
        float a, s1, s2;
        hpv_act<ACT>(z.v.a[j], a, s1, s2);
        float zb = g.v.a[j] * s1;
        if constexpr (M::DX) {
            float dz = z.dx.a[j];
            zb = fmaf(g.dx.a[j] * s2, dz, zb);
            float dzb = g.dx.a[j] * s1;
            if constexpr (M::EX) {
                float s3 = hpv_act_s3<ACT>(a, s1);
                float d2z = z.ex.a[j];
                zb = fmaf(g.ex.a[j], fmaf(s3 * dz, dz, s2 * d2z), zb);
                dzb = fmaf(2.0f * g.ex.a[j] * s2, dz, dzb);
                z.ex.a[j] = fmaf(s2 * dz, dz, s1 * d2z);
                g.ex.a[j] = g.ex.a[j] * s1;
            }
            z.dx.a[j] = s1 * dz;
            g.dx.a[j] = dzb;
        }
        if constexpr (M::DY) {
            float dz = z.dy.a[j];
            zb = fmaf(g.dy.a[j] * s2, dz, zb);
            float dzb = g.dy.a[j] * s1;
            if constexpr (M::EY) {
                float s3 = hpv_act_s3<ACT>(a, s1);
                float d2z = z.ey.a[j];
                zb = fmaf(g.ey.a[j], fmaf(s3 * dz, dz, s2 * d2z), zb);
                dzb = fmaf(2.0f * g.ey.a[j] * s2, dz, dzb);
                z.ey.a[j] = fmaf(s2 * dz, dz, s1 * d2z);
                g.ey.a[j] = g.ey.a[j] * s1;
            }
            z.dy.a[j] = s1 * dz;
            g.dy.a[j] = dzb;
        }
        z.v.a[j] = a;
        g.v.a[j] = zb;
    }
}
