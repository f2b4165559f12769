// Per-point MLP arithmetic of the hp-VPINN hot path: net_u (P2D:158-173, P1D:128-142, ADI:219-234) and its
// input derivatives (tf.gradients chains of P2D:175-185, P1D:144-148, ADI:236-245) in analytic forward mode,
// plus the hand-derived reverse sweep through that forward-mode computation (what TF builds for
// AdamOptimizer.minimize, P2D:131-132).  fp32, FFMA.  Compiles for the device (nvcc) and, for the
// thread-emulated kernel tests under tests/emu, for the host (g++).
#pragma once
#include <math.h>
#include "hpv_types.h"
#include "hpv_pair.cuh"

enum { HPV_ACT_SIN = 0, HPV_ACT_TANH = 1 };

struct alignas(16) HpvF4 { float x, y, z, w; };

HPV_HD HpvF4 hpv_ld4(const float* p) { return *reinterpret_cast<const HpvF4*>(p); }
HPV_HD void hpv_st4(float* p, const HpvF4& v) { *reinterpret_cast<HpvF4*>(p) = v; }

// ---- padded parameter layout --------------------------------------------------------------------------
// [W1: DIM x HP][b1: HP] { [W_l: HP x HP][b_l: HP][W_l^T: HP x HP] } x (nhid-1) [Wo: HP][bo, 0, 0, 0]
// The transposed copy of each hidden matrix serves the adjoint product of the reverse sweep
// (out[i] = sum_j in[j] W[i][j]), which then runs as the same broadcast-input / uniform-weight FFMA2 loop as
// the forward product.  Hidden-layer offsets depend on the layer index only (they stay in uniform registers).
// Hidden width H is zero-padded to HP: a padded unit has z = 0 -> sigma(0) = 0 for both sin and tanh and
// all its tangents are 0, so the padding is exact.
HPV_HD int hpv_off_w1() { return 0; }
HPV_HD int hpv_off_b1(int dim, int hp) { return dim * hp; }
HPV_HD int hpv_off_wl(int dim, int hp, int l /*1..nhid-1*/) { return (dim + 1) * hp + (l - 1) * (2 * hp * hp + hp); }
HPV_HD int hpv_off_wt(int dim, int hp, int l /*1..nhid-1*/) { return hpv_off_wl(dim, hp, l) + hp * hp + hp; }
HPV_HD int hpv_off_wo(int dim, int hp, int nhid) { return (dim + 1) * hp + (nhid - 1) * (2 * hp * hp + hp); }
HPV_HD int hpv_theta_pad_n(int dim, int hp, int nhid) { return hpv_off_wo(dim, hp, nhid) + hp + 4; }

// ---- activations ----------------------------------------------------------------------------------------
// tanh as 1 - 2/(exp(2z)+1): absolute error ~1e-7 (what matters: the value feeds dot products), saturates
// cleanly to +-1, no branches.
HPV_HD float hpv_tanh(float z) {
    const float e = hpv_ex2(z * 2.8853900817779268f);    // exp(2z): MUFU.EX2, then MUFU.RCP
    return fmaf(-2.0f, hpv_rcp(e + 1.0f), 1.0f);
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
    // <2, 1, 0> is the DIRECTIONAL mode of the reverse sweep (never a forward mode, see hpv_canon_mode): the one
    // tangent channel carries the derivative along a per-point direction (vx, vy).  For forms that use first
    // derivatives only, sum_p (gx u_x + gy u_y)(p) = sum_p D_{(gx,gy)(p)} u(p), so the gradient of the loss needs
    // the value and ONE tangent per point instead of two (2/3 of the arithmetic of mode <2, 1, 1>).
    static constexpr bool DIR = (DIM == 2) && (MX == 1) && (MY == 0);
};

// A channel of one layer at one point: HP units held as HP/2 packed pairs (units 2m, 2m+1).
template <int HP, bool ON>
struct HpvVec { hpv_pair p[ON ? HP / 2 : 1]; };

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
    const hpv_pair xx = hpv_dup(x), yy = hpv_dup(y);
#pragma unroll
    for (int m = 0; m < HP / 2; ++m) {
        const hpv_pair wx = hpv_pack(W1[2 * m], W1[2 * m + 1]);
        hpv_pair zz = hpv_fma2r(xx, wx, hpv_pack(b1[2 * m], b1[2 * m + 1]));
        if constexpr (DIM == 2) {
            const hpv_pair wy = hpv_pack(W1[HP + 2 * m], W1[HP + 2 * m + 1]);
            zz = hpv_fma2r(yy, wy, zz);
            if constexpr (M::DY) z.dy.p[m] = wy;
        }
        z.v.p[m] = zz;
        if constexpr (M::DX) z.dx.p[m] = wx;
        if constexpr (M::EX) z.ex.p[m] = hpv_dup(0.0f);
        if constexpr (M::EY) z.ey.p[m] = hpv_dup(0.0f);
    }
}

// First layer of the directional mode: z as above, tangent seed dz = vx W1[0,:] + vy W1[1,:].
template <int HP>
HPV_HD void hpv_layer1_pre_dir(const float* th, float x, float y, float vx, float vy, HpvState<2, 1, 0, HP>& z) {
    const float* W1 = th + hpv_off_w1();
    const float* b1 = th + hpv_off_b1(2, HP);
    const hpv_pair xx = hpv_dup(x), yy = hpv_dup(y), vxx = hpv_dup(vx), vyy = hpv_dup(vy);
#pragma unroll
    for (int m = 0; m < HP / 2; ++m) {
        const hpv_pair wx = hpv_pack(W1[2 * m], W1[2 * m + 1]);
        const hpv_pair wy = hpv_pack(W1[HP + 2 * m], W1[HP + 2 * m + 1]);
        z.v.p[m] = hpv_fma2r(yy, wy, hpv_fma2r(xx, wx, hpv_pack(b1[2 * m], b1[2 * m + 1])));
        z.dx.p[m] = hpv_fma2r(vyy, wy, hpv_mul2(vxx, wx));
    }
}

// tanh of a pair and its first derivative factor: a = 1 - 2/(exp(2z)+1), s1 = 1 - a^2.
HPV_HD void hpv_tanh2(hpv_pair z, hpv_pair& a, hpv_pair& s1) {
    const hpv_pair t = hpv_mul2(z, hpv_dup(2.8853900817779268f));
    float t0, t1;
    hpv_unpack(t, t0, t1);
    const hpv_pair d = hpv_add2(hpv_pack(hpv_ex2(t0), hpv_ex2(t1)), hpv_dup(1.0f));
    float d0, d1;
    hpv_unpack(d, d0, d1);
    a = hpv_fma2r(hpv_pack(hpv_rcp(d0), hpv_rcp(d1)), hpv_dup(-2.0f), hpv_dup(1.0f));
    s1 = hpv_fma2r(hpv_mul2(a, a), hpv_dup(-1.0f), hpv_dup(1.0f));
}

// Activation value and its first three derivatives for a pair of units.
template <int ACT>
HPV_HD void hpv_act2(hpv_pair z, hpv_pair& a, hpv_pair& s1, hpv_pair& s2, hpv_pair& s3, bool need3) {
    if constexpr (ACT == HPV_ACT_TANH) {
        hpv_tanh2(z, a, s1);
        s2 = hpv_mul2(hpv_mul2(a, s1), hpv_dup(-2.0f));
        if (need3) s3 = hpv_mul2(hpv_mul2(s1, hpv_fma2r(hpv_mul2(a, a), hpv_dup(-3.0f), hpv_dup(1.0f))), hpv_dup(-2.0f));
    } else {
        float z0, z1, a0, a1, b0, b1, c0, c1;
        hpv_unpack(z, z0, z1);
        hpv_act<ACT>(z0, a0, b0, c0);
        hpv_act<ACT>(z1, a1, b1, c1);
        a = hpv_pack(a0, a1); s1 = hpv_pack(b0, b1); s2 = hpv_pack(c0, c1);
        if (need3) s3 = hpv_pack(hpv_act_s3<ACT>(a0, b0), hpv_act_s3<ACT>(a1, b1));
    }
}

// "Mixed" state of a layer, the form the reverse sweep keeps: the tangent channels as pre-activations (dz, d2z),
// the value channel as the pre-activation z for sin, but as the POST-activation a = tanh z for tanh -- every
// derivative of tanh is a polynomial in a (s1 = 1 - a^2, s2 = -2 a s1, s3 = -2 s1 (1 - 3 a^2)), so whatever is
// derived from a mixed state later costs no second exp/reciprocal.
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_to_mixed(HpvState<DIM, MX, MY, HP>& s) {
    if constexpr (ACT == HPV_ACT_TANH) {
#pragma unroll
        for (int m = 0; m < HP / 2; ++m) {
            hpv_pair a, s1;
            hpv_tanh2(s.v.p[m], a, s1);
            s.v.p[m] = a;
        }
    }
}

// Activation value and derivatives of a pair of units from the value channel of a plain (MIXED = false:
// pre-activation) or mixed state.
template <int ACT, bool MIXED>
HPV_HD void hpv_act2m(hpv_pair zv, hpv_pair& a, hpv_pair& s1, hpv_pair& s2, hpv_pair& s3, bool need3) {
    if constexpr (MIXED && ACT == HPV_ACT_TANH) {
        a = zv;
        s1 = hpv_fma2r(hpv_mul2(a, a), hpv_dup(-1.0f), hpv_dup(1.0f));
        s2 = hpv_mul2(hpv_mul2(a, s1), hpv_dup(-2.0f));
        if (need3) s3 = hpv_mul2(hpv_mul2(s1, hpv_fma2r(hpv_mul2(a, a), hpv_dup(-3.0f), hpv_dup(1.0f))), hpv_dup(-2.0f));
    } else {
        hpv_act2<ACT>(zv, a, s1, s2, s3, need3);
    }
}

// One pair of units m of hpv_activate / hpv_activate_bwd (below), results in o.  The fused compute-and-store
// helpers of hpv_slot.cuh use them quad by quad, so that a stored state never has to exist as a whole in
// registers (which would have to be gathered into aligned quads for the 128-bit stores, one MOV per register).
template <int DIM, int MX, int MY>
struct HpvPairOut { hpv_pair v, dx, dy, ex, ey; };

//   h = s(z);  dh = s'(z) dz;  d2h = s''(z) dz^2 + s'(z) d2z
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED>
HPV_HD void hpv_activate_pair(const HpvState<DIM, MX, MY, HP>& s, int m, HpvPairOut<DIM, MX, MY>& o) {
    typedef HpvMode<DIM, MX, MY> M;
    hpv_pair a, s1, s2, s3;
    hpv_act2m<ACT, MIXED>(s.v.p[m], a, s1, s2, s3, false);
    o.v = a;
    if constexpr (M::DX) {
        const hpv_pair dz = s.dx.p[m];
        if constexpr (M::EX) o.ex = hpv_fma2r(hpv_mul2(s2, dz), dz, hpv_mul2(s1, s.ex.p[m]));
        o.dx = hpv_mul2(s1, dz);
    }
    if constexpr (M::DY) {
        const hpv_pair dz = s.dy.p[m];
        if constexpr (M::EY) o.ey = hpv_fma2r(hpv_mul2(s2, dz), dz, hpv_mul2(s1, s.ey.p[m]));
        o.dy = hpv_mul2(s1, dz);
    }
}

// Reverse of the above for pair m.  In: z (pre-activations or a mixed state) and the adjoints g of the
// post-activations.  Out: adjoints of the pre-activations.
//   zbar   = hbar s1 + sum_d [ dhbar_d s2 dz_d + d2hbar_d (s3 dz_d^2 + s2 d2z_d) ]
//   dzbar  = dhbar s1 + 2 d2hbar s2 dz
//   d2zbar = d2hbar s1
// with s1, s2, s3 the first three derivatives of the activation at z (tanh: s2 = -2 a s1, s3 = -2 s1 (1 - 3 a^2)).
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED>
HPV_HD void hpv_activate_bwd_pair(const HpvState<DIM, MX, MY, HP>& z, const HpvState<DIM, MX, MY, HP>& g, int m,
                                  HpvPairOut<DIM, MX, MY>& o) {
    typedef HpvMode<DIM, MX, MY> M;
    hpv_pair a, s1, s2, s3;
    hpv_act2m<ACT, MIXED>(z.v.p[m], a, s1, s2, s3, M::EX || M::EY);
    hpv_pair zb = hpv_mul2(g.v.p[m], s1);
    if constexpr (M::DX) {
        const hpv_pair dz = z.dx.p[m], gd = g.dx.p[m];
        zb = hpv_fma2r(hpv_mul2(gd, s2), dz, zb);
        hpv_pair dzb = hpv_mul2(gd, s1);
        if constexpr (M::EX) {
            const hpv_pair d2z = z.ex.p[m], ge = g.ex.p[m];
            zb = hpv_fma2r(ge, hpv_fma2r(hpv_mul2(s3, dz), dz, hpv_mul2(s2, d2z)), zb);
            dzb = hpv_fma2r(hpv_mul2(hpv_mul2(ge, s2), hpv_dup(2.0f)), dz, dzb);
            o.ex = hpv_mul2(ge, s1);
        }
        o.dx = dzb;
    }
    if constexpr (M::DY) {
        const hpv_pair dz = z.dy.p[m], gd = g.dy.p[m];
        zb = hpv_fma2r(hpv_mul2(gd, s2), dz, zb);
        hpv_pair dzb = hpv_mul2(gd, s1);
        if constexpr (M::EY) {
            const hpv_pair d2z = z.ey.p[m], ge = g.ey.p[m];
            zb = hpv_fma2r(ge, hpv_fma2r(hpv_mul2(s3, dz), dz, hpv_mul2(s2, d2z)), zb);
            dzb = hpv_fma2r(hpv_mul2(hpv_mul2(ge, s2), hpv_dup(2.0f)), dz, dzb);
            o.ey = hpv_mul2(ge, s1);
        }
        o.dy = dzb;
    }
    o.v = zb;
}

template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_put_pair(HpvState<DIM, MX, MY, HP>& s, int m, const HpvPairOut<DIM, MX, MY>& o) {
    typedef HpvMode<DIM, MX, MY> M;
    s.v.p[m] = o.v;
    if constexpr (M::DX) s.dx.p[m] = o.dx;
    if constexpr (M::DY) s.dy.p[m] = o.dy;
    if constexpr (M::EX) s.ex.p[m] = o.ex;
    if constexpr (M::EY) s.ey.p[m] = o.ey;
}

// Pre-activations (z, dz, d2z) [or a mixed state] -> post-activations (h, dh, d2h), in place.
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED = false>
HPV_HD void hpv_activate(HpvState<DIM, MX, MY, HP>& s) {
#pragma unroll
    for (int m = 0; m < HP / 2; ++m) {
        HpvPairOut<DIM, MX, MY> o;
        hpv_activate_pair<DIM, MX, MY, HP, ACT, MIXED>(s, m, o);
        hpv_put_pair<DIM, MX, MY, HP>(s, m, o);
    }
}

// Reverse of hpv_activate, in place in g (z: pre-activations, or a mixed state with MIXED).
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED = false>
HPV_HD void hpv_activate_bwd(const HpvState<DIM, MX, MY, HP>& z, HpvState<DIM, MX, MY, HP>& g) {
#pragma unroll
    for (int m = 0; m < HP / 2; ++m) {
        HpvPairOut<DIM, MX, MY> o;
        hpv_activate_bwd_pair<DIM, MX, MY, HP, ACT, MIXED>(z, g, m, o);
        hpv_put_pair<DIM, MX, MY, HP>(g, m, o);
    }
}
