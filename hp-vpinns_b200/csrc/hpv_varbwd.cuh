// Backward of the variational loss (what tf.train.AdamOptimizer.minimize differentiates, P2D:131-132,
// P1D:102-104, ADI:191-193), as two kernels:
//   K1  hpv_adjproj_body : adjoint of the projection.  Rbar = dloss/dU = 2 Res / (Ntx Nty) per element, then
//       Gbar_t[j][i] = c_t * sum_k sum_r L_t[j][k] Rbar[k][r] R_t[i][r]  -> adjoint of the point field of term t.
//   K2  hpv_mlpbwd_body  : per quadrature point, recompute the forward-mode MLP, run the reverse sweep through
//       it, and accumulate the weight gradients as small GEMMs over the points of a tile (deterministic).
// A third, tiny kernel (hpv_gradreduce_body) sums the per-CTA partial gradients in a fixed order.
#pragma once
#include "hpv_varfwd.cuh"

#define HPV_ADJ_RS 16          // rows of an element handled by one CTA of K1

struct HpvAdjSmem { int tab[HPV_NTAB], Rbar, V, total; };

HPV_HD HpvAdjSmem hpv_adj_smem(const HpvVarArgs& a) {
    HpvAdjSmem s;
    int o = 0;
    int m = hpv_tab_mask(a);
    for (int t = 0; t < HPV_NTAB; ++t) {
        s.tab[t] = -1;
        if (m & (1 << t)) { s.tab[t] = o; o += a.Q * HPV_NP; }
    }
    s.Rbar = o; o += HPV_NP * HPV_NP;
    s.V = o; o += a.n_terms * HPV_ADJ_RS * HPV_NP;
    s.total = o;
    return s;
}

// Gbar layout: [term][element][rows*Q]
struct HpvAdjArgs {
    HpvVarArgs v;
    float* Gbar;
    int slabs_per_el;
};

HPV_HD void hpv_adjproj_body(const HpvCta& c, const HpvAdjArgs& aa) {
    const HpvVarArgs& a = aa.v;
    const HpvAdjSmem L = hpv_adj_smem(a);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_R = sm + L.Rbar;
    float* s_V = sm + L.V;
    const int T = c.nthreads, tid = c.tid, Q = a.Q;
    const int e = c.bid / aa.slabs_per_el, slab = c.bid - e * aa.slabs_per_el;
    const int j0 = slab * HPV_ADJ_RS;
    int nrows = a.rows - j0;
    if (nrows > HPV_ADJ_RS) nrows = HPV_ADJ_RS;
    const int npts_el = a.rows * Q;

    for (int t = 0; t < HPV_NTAB; ++t) {
        if (L.tab[t] < 0) continue;
        const float* src = a.tab[t];
        float* dst = sm + L.tab[t];
        for (int i = tid * 4; i < Q * HPV_NP; i += T * 4) hpv_st4(dst + i, hpv_ld4(src + i));
    }
    const int ntx_e = a.el_ntest[2 * e + 0], nty_e = a.el_ntest[2 * e + 1];
    const float rs = a.loss_scale * 2.0f / (float)(ntx_e * nty_e);
    for (int idx = tid; idx < HPV_NP * HPV_NP; idx += T) {
        const int k = idx >> 6, r = idx & 63;
        float v = 0.0f;
        if (k < nty_e && r < ntx_e) v = rs * a.Res[((size_t)e * a.nty + k) * a.ntx + r];
        s_R[idx] = v;
    }
    hpv_sync(c);

    // V_t[jl][r] = sum_k L_t[j0+jl][k] * Rbar[k][r]
    {
        const int nitems = a.n_terms * nrows * (HPV_NP / 4);
        for (int item = tid; item < nitems; item += T) {
            const int r4 = item & 15, rest = item >> 4;
            const int jl = rest % nrows, t = rest / nrows;
            const float* Lt = sm + L.tab[a.terms[t].ltab] + (j0 + jl) * HPV_NP;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            for (int k = 0; k < nty_e; ++k) {
                const float lv = Lt[k];
                const HpvF4 w = hpv_ld4(s_R + k * HPV_NP + 4 * r4);
                a0 = fmaf(lv, w.x, a0); a1 = fmaf(lv, w.y, a1); a2 = fmaf(lv, w.z, a2); a3 = fmaf(lv, w.w, a3);
            }
            HpvF4 o; o.x = a0; o.y = a1; o.z = a2; o.w = a3;
            hpv_st4(s_V + (t * HPV_ADJ_RS + jl) * HPV_NP + 4 * r4, o);
        }
    }
    hpv_sync(c);

    // Gbar_t[j][i] = c_t * sum_r V_t[jl][r] * R_t[i][r]
    {
        const float hwx = a.el_geom[4 * e + 1], hwy = a.el_geom[4 * e + 3];
        const int nitems = a.n_terms * nrows * Q;
        for (int item = tid; item < nitems; item += T) {
            const int i = item % Q, rest = item / Q;
            const int jl = rest % nrows, t = rest / nrows;
            const float* Vt = s_V + (t * HPV_ADJ_RS + jl) * HPV_NP;
            const float* R = sm + L.tab[a.terms[t].rtab] + i * HPV_NP;
            float acc = 0.0f;
            for (int r4 = 0; r4 < HPV_NP / 4; ++r4) {
                const HpvF4 v = hpv_ld4(Vt + 4 * r4), w = hpv_ld4(R + 4 * r4);
                acc = fmaf(v.x, w.x, acc); acc = fmaf(v.y, w.y, acc); acc = fmaf(v.z, w.z, acc); acc = fmaf(v.w, w.w, acc);
            }
            const float ct = hpv_term_scale(a.terms[t], hwx, hwy);
            aa.Gbar[((size_t)t * a.n_el + e) * npts_el + (j0 + jl) * Q + i] = ct * acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2: reverse sweep through the MLP.
// ---------------------------------------------------------------------------------------------------------
template <int HP> struct HpvSP { static constexpr int value = ((HP / 4) % 2 == 1) ? HP : HP + 4; };

struct HpvBwdArgs {
    HpvVarArgs v;
    const float* Gbar;         // [term][n_points]  (n_points = n_el*rows*Q for the variational loss)
    int n_points;
    int n_tiles;               // tiles of blockDim points
    // scattered-point mode (boundary / PINN losses): coordinates and per-point adjoints are given directly
    const float* pts;          // [n][dim] or null (then the points are the element quadrature points)
};

template <int DIM, int MX, int MY, int HP>
struct HpvBwdSmem {
    typedef HpvMode<DIM, MX, MY> M;
    static constexpr int SP = HpvSP<HP>::value;
    int th, gw, slots, in0, go, scratch, red, total, NS, slot_sz;
    HPV_HD HpvBwdSmem(int theta_pad_n, int nhid, int T) {
        int o = 0;
        th = o; o += hpv_align4(theta_pad_n);
        gw = o; o += hpv_align4(theta_pad_n);
        NS = nhid - 1 > 2 ? nhid - 1 : 2;
        slot_sz = M::NCH * T * SP;
        slots = o; o += NS * slot_sz;
        in0 = o; o += 3 * T * 4;
        go = o; o += M::NCH * T * 4;
        scratch = o; o += T * 16;
        red = o; o += T;
        total = o;
    }
};

template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_store_state(float* slot, int T, int tid, const HpvState<DIM, MX, MY, HP>& s) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value;
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvF4 o;
        o.x = s.v.a[4 * j4]; o.y = s.v.a[4 * j4 + 1]; o.z = s.v.a[4 * j4 + 2]; o.w = s.v.a[4 * j4 + 3];
        hpv_st4(slot + (M::C_V * T + tid) * SP + 4 * j4, o);
        if constexpr (M::DX) {
            o.x = s.dx.a[4 * j4]; o.y = s.dx.a[4 * j4 + 1]; o.z = s.dx.a[4 * j4 + 2]; o.w = s.dx.a[4 * j4 + 3];
            hpv_st4(slot + (M::C_DX * T + tid) * SP + 4 * j4, o);
        }
        if constexpr (M::DY) {
            o.x = s.dy.a[4 * j4]; o.y = s.dy.a[4 * j4 + 1]; o.z = s.dy.a[4 * j4 + 2]; o.w = s.dy.a[4 * j4 + 3];
            hpv_st4(slot + (M::C_DY * T + tid) * SP + 4 * j4, o);
        }
        if constexpr (M::EX) {
            o.x = s.ex.a[4 * j4]; o.y = s.ex.a[4 * j4 + 1]; o.z = s.ex.a[4 * j4 + 2]; o.w = s.ex.a[4 * j4 + 3];
            hpv_st4(slot + (M::C_EX * T + tid) * SP + 4 * j4, o);
        }
        if constexpr (M::EY) {
            o.x = s.ey.a[4 * j4]; o.y = s.ey.a[4 * j4 + 1]; o.z = s.ey.a[4 * j4 + 2]; o.w = s.ey.a[4 * j4 + 3];
            hpv_st4(slot + (M::C_EY * T + tid) * SP + 4 * j4, o);
        }
    }
}

template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_load_state(const float* slot, int T, int tid, HpvState<DIM, MX, MY, HP>& s) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value;
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvF4 o = hpv_ld4(slot + (M::C_V * T + tid) * SP + 4 * j4);
        s.v.a[4 * j4] = o.x; s.v.a[4 * j4 + 1] = o.y; s.v.a[4 * j4 + 2] = o.z; s.v.a[4 * j4 + 3] = o.w;
        if constexpr (M::DX) {
            o = hpv_ld4(slot + (M::C_DX * T + tid) * SP + 4 * j4);
            s.dx.a[4 * j4] = o.x; s.dx.a[4 * j4 + 1] = o.y; s.dx.a[4 * j4 + 2] = o.z; s.dx.a[4 * j4 + 3] = o.w;
        }
        if constexpr (M::DY) {
            o = hpv_ld4(slot + (M::C_DY * T + tid) * SP + 4 * j4);
            s.dy.a[4 * j4] = o.x; s.dy.a[4 * j4 + 1] = o.y; s.dy.a[4 * j4 + 2] = o.z; s.dy.a[4 * j4 + 3] = o.w;
        }
        if constexpr (M::EX) {
            o = hpv_ld4(slot + (M::C_EX * T + tid) * SP + 4 * j4);
            s.ex.a[4 * j4] = o.x; s.ex.a[4 * j4 + 1] = o.y; s.ex.a[4 * j4 + 2] = o.z; s.ex.a[4 * j4 + 3] = o.w;
        }
        if constexpr (M::EY) {
            o = hpv_ld4(slot + (M::C_EY * T + tid) * SP + 4 * j4);
            s.ey.a[4 * j4] = o.x; s.ey.a[4 * j4 + 1] = o.y; s.ey.a[4 * j4 + 2] = o.z; s.ey.a[4 * j4 + 3] = o.w;
        }
    }
}

// Weight-gradient GEMM over the T points of a tile:  D[i][j] += sum_ch sum_p IN[ch][p][i] * ADJ[ch][p][j].
// 4x4 output tiles, K (points) split over KS threads per tile, partials combined through shared memory in a
// fixed order (deterministic).  `bias` adds a virtual row of ones in the value channel (-> bias gradient).
// dst_kind: 0 hidden layer (dst = W[HP][HP], dstb = b), 1 output layer (dst = Wo[HP], dstb = bo),
//           2 first layer (IN rows = x, y, 1: dst = W1[DIM][HP], dstb = b1).
HPV_HD void hpv_wgrad_gemm(const HpvCta& c, const float* IN, int spi, const float* ADJ, int spa, int nch,
                           int NI, int NJ, bool bias, int dst_kind, float* dst, float* dstb, int hp, int dim,
                           float* scratch) {
    const int T = c.nthreads, tid = c.tid;
    const int ntiles = (NI + (bias ? 1 : 0)) * NJ;
    int KS = 1;
    while (KS * 2 * ntiles <= T) KS *= 2;
    const int tau = tid / KS, ks = tid - tau * KS;
    const bool active = tau < ntiles;
    const int it = active ? tau / NJ : 0, jt = active ? tau - it * NJ : 0;
    const bool brow = bias && it == NI;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    if (active) {
        for (int ch = 0; ch < nch; ++ch) {
            if (brow && ch != 0) break;
            const float* inb = IN + (size_t)ch * T * spi + 4 * it;
            const float* adb = ADJ + (size_t)ch * T * spa + 4 * jt;
            for (int p = ks; p < T; p += KS) {
                HpvF4 a4;
                if (brow) { a4.x = 1.0f; a4.y = a4.z = a4.w = 0.0f; }
                else a4 = hpv_ld4(inb + p * spi);
                const HpvF4 b4 = hpv_ld4(adb + p * spa);
                const float as[4] = {a4.x, a4.y, a4.z, a4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(as[i], bs[j], acc[i][j]);
            }
        }
    }
    if (KS > 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            HpvF4 o; o.x = acc[i][0]; o.y = acc[i][1]; o.z = acc[i][2]; o.w = acc[i][3];
            hpv_st4(scratch + tid * 16 + 4 * i, o);
        }
        hpv_sync(c);
        if (active && ks == 0) {
            for (int s = 1; s < KS; ++s) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const HpvF4 v = hpv_ld4(scratch + (tid + s) * 16 + 4 * i);
                    acc[i][0] += v.x; acc[i][1] += v.y; acc[i][2] += v.z; acc[i][3] += v.w;
                }
            }
        }
    }
    if (active && ks == 0) {
        if (dst_kind == 0) {
            if (brow) { for (int j = 0; j < 4; ++j) dstb[4 * jt + j] += acc[0][j]; }
            else for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dst[(4 * it + i) * hp + 4 * jt + j] += acc[i][j];
        } else if (dst_kind == 1) {
            if (brow) dstb[0] += acc[0][0];
            else for (int i = 0; i < 4; ++i) dst[4 * it + i] += acc[i][0];
        } else {
            for (int j = 0; j < 4; ++j) {
                dst[4 * jt + j] += acc[0][j];
                if (dim == 2) dst[hp + 4 * jt + j] += acc[1][j];
                dstb[4 * jt + j] += acc[2][j];
            }
        }
    }
}

template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_mlpbwd_body(const HpvCta& c, const HpvBwdArgs& ba) {
    typedef HpvMode<DIM, MX, MY> M;
    typedef HpvState<DIM, MX, MY, HP> State;
    constexpr int SP = HpvSP<HP>::value;
    const HpvVarArgs& a = ba.v;
    const int T = c.nthreads, tid = c.tid, nhid = a.nhid, top = nhid - 1;
    const HpvBwdSmem<DIM, MX, MY, HP> L(a.theta_pad_n, nhid, T);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_th = sm + L.th;
    float* s_gw = sm + L.gw;
    float* s_in0 = sm + L.in0;
    float* s_go = sm + L.go;
    float* s_scr = sm + L.scratch;
    float* s_red = sm + L.red;
#define HPV_SLOT(l) (sm + L.slots + ((l) == 0 ? 1 : (l) - 1) * L.slot_sz)

    for (int i = tid; i < a.theta_pad_n; i += T) { s_th[i] = a.theta_pad[i]; s_gw[i] = 0.0f; }
    const float eps = a.eps[0];
    float coef[HPV_MAX_TERMS][HPV_NFIELDS], coef1[HPV_MAX_TERMS][HPV_NFIELDS];
    for (int t = 0; t < HPV_MAX_TERMS; ++t)
        for (int f = 0; f < HPV_NFIELDS; ++f) {
            coef[t][f] = (t < a.n_terms) ? fmaf(eps, a.terms[t].a1[f], a.terms[t].a0[f]) : 0.0f;
            coef1[t][f] = (t < a.n_terms) ? a.terms[t].a1[f] : 0.0f;
        }
    hpv_sync(c);

    const int Q = a.Q, npts_el = a.rows * Q;
    float deps = 0.0f;
    const int tile_begin = (int)(((long long)c.bid * ba.n_tiles) / c.nblocks);
    const int tile_end = (int)(((long long)(c.bid + 1) * ba.n_tiles) / c.nblocks);
    const float* Wo = s_th + hpv_off_wo(DIM, HP, nhid);

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int gp = tile * T + tid;
        const bool valid = gp < ba.n_points;
        float x = 0.0f, y = 0.0f;
        float gbar[HPV_MAX_TERMS] = {0.0f, 0.0f};
        if (valid) {
            if (ba.pts) {
                x = ba.pts[(size_t)gp * DIM];
                if (DIM == 2) y = ba.pts[(size_t)gp * DIM + 1];
            } else {
                const int e = gp / npts_el, p = gp - e * npts_el;
                const int j = p / Q, i = p - j * Q;
                x = fmaf(a.el_geom[4 * e + 1], a.xi1[i], a.el_geom[4 * e + 0]);
                if (DIM == 2) y = fmaf(a.el_geom[4 * e + 3], a.xi1[j], a.el_geom[4 * e + 2]);
            }
            for (int t = 0; t < a.n_terms; ++t) gbar[t] = ba.Gbar[(size_t)t * ba.n_points + gp];
        }

        // ---- forward recompute, keeping the hidden pre-activations of layers 1..top-1 in shared memory ----
        State pre, h, g;
        hpv_layer1_pre<DIM, MX, MY, HP>(s_th, x, y, pre);
        for (int l = 1; l <= top; ++l) {
            h = pre;
            hpv_activate<DIM, MX, MY, HP, ACT>(h);
            const float* W = s_th + hpv_off_wl(DIM, HP, l);
            hpv_matmul<DIM, MX, MY, HP>(W, W + HP * HP, h, pre);
            if (l < top) hpv_store_state<DIM, MX, MY, HP>(HPV_SLOT(l), T, tid, pre);
        }
        h = pre;
        hpv_activate<DIM, MX, MY, HP, ACT>(h);                  // h_top
        float f[HPV_NFIELDS], gf[HPV_NFIELDS];
        hpv_output<DIM, MX, MY, HP>(Wo, h, f);
#pragma unroll
        for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = 0.0f;
        for (int t = 0; t < a.n_terms; ++t) {
            float d1 = 0.0f;
#pragma unroll
            for (int k = 0; k < HPV_NFIELDS; ++k) {
                gf[k] = fmaf(gbar[t], coef[t][k], gf[k]);
                d1 = fmaf(coef1[t][k], f[k], d1);
            }
            deps = fmaf(gbar[t], d1, deps);
        }

        // ---- output layer: Wo/bo gradient, adjoint of h_top ----
        hpv_store_state<DIM, MX, MY, HP>(HPV_SLOT(top), T, tid, h);
        {
            HpvF4 o; o.y = o.z = o.w = 0.0f;
            o.x = gf[0]; hpv_st4(s_go + (M::C_V * T + tid) * 4, o);
            if constexpr (M::DX) { o.x = gf[1]; hpv_st4(s_go + (M::C_DX * T + tid) * 4, o); }
            if constexpr (M::DY) { o.x = gf[2]; hpv_st4(s_go + (M::C_DY * T + tid) * 4, o); }
            if constexpr (M::EX) { o.x = gf[3]; hpv_st4(s_go + (M::C_EX * T + tid) * 4, o); }
            if constexpr (M::EY) { o.x = gf[4]; hpv_st4(s_go + (M::C_EY * T + tid) * 4, o); }
        }
#pragma unroll
        for (int j = 0; j < HP; ++j) {
            const float w = Wo[j];
            g.v.a[j] = gf[0] * w;
            if constexpr (M::DX) g.dx.a[j] = gf[1] * w;
            if constexpr (M::DY) g.dy.a[j] = gf[2] * w;
            if constexpr (M::EX) g.ex.a[j] = gf[3] * w;
            if constexpr (M::EY) g.ey.a[j] = gf[4] * w;
        }
        hpv_sync(c);
        hpv_wgrad_gemm(c, HPV_SLOT(top), SP, s_go, 4, M::NCH, HP / 4, 1, true, 1,
                       s_gw + hpv_off_wo(DIM, HP, nhid), s_gw + hpv_off_wo(DIM, HP, nhid) + HP, HP, DIM, s_scr);
        hpv_sync(c);

        // ---- hidden layers, top down ----
        for (int l = top; l >= 1; --l) {
            hpv_activate_bwd<DIM, MX, MY, HP, ACT>(pre, g);     // g := adjoint of the pre-activations of layer l
            hpv_store_state<DIM, MX, MY, HP>(HPV_SLOT(l), T, tid, g);
            const float* W = s_th + hpv_off_wl(DIM, HP, l);
            State gn;
            hpv_matmul_t<DIM, MX, MY, HP>(W, g, gn);             // adjoint of h_{l-1}
            if (l - 1 >= 1) hpv_load_state<DIM, MX, MY, HP>(HPV_SLOT(l - 1), T, tid, pre);
            else hpv_layer1_pre<DIM, MX, MY, HP>(s_th, x, y, pre);
            h = pre;
            hpv_activate<DIM, MX, MY, HP, ACT>(h);              // h_{l-1}: left factor of the W_l gradient
            hpv_store_state<DIM, MX, MY, HP>(HPV_SLOT(l - 1), T, tid, h);
            hpv_sync(c);
            float* gW = s_gw + hpv_off_wl(DIM, HP, l);
            hpv_wgrad_gemm(c, HPV_SLOT(l - 1), SP, HPV_SLOT(l), SP, M::NCH, HP / 4, HP / 4, true, 0,
                           gW, gW + HP * HP, HP, DIM, s_scr);
            hpv_sync(c);
            g = gn;
        }

        // ---- first layer ----
        hpv_activate_bwd<DIM, MX, MY, HP, ACT>(pre, g);
        hpv_store_state<DIM, MX, MY, HP>(HPV_SLOT(0), T, tid, g);
        {
            HpvF4 o;
            o.x = x; o.y = y; o.z = 1.0f; o.w = 0.0f; hpv_st4(s_in0 + (0 * T + tid) * 4, o);
            o.x = 1.0f; o.y = 0.0f; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (1 * T + tid) * 4, o);
            o.x = 0.0f; o.y = 1.0f; hpv_st4(s_in0 + (2 * T + tid) * 4, o);
        }
        hpv_sync(c);
        // channels that feed W1: value (x, y, 1), d/dx (1, 0, 0), d/dy (0, 1, 0); second-derivative seeds are 0.
        // The IN rows are ordered v, dx, dy, which is also the order of the stored channels (C_V, C_DX, C_DY).
        {
            float* gW1 = s_gw + hpv_off_w1();
            float* gb1 = s_gw + hpv_off_b1(DIM, HP);
            constexpr int nch1 = 1 + (M::DX ? 1 : 0) + (M::DY ? 1 : 0);
            hpv_wgrad_gemm(c, s_in0, 4, HPV_SLOT(0), SP, nch1, 1, HP / 4, false, 2, gW1, gb1, HP, DIM, s_scr);
            hpv_sync(c);
        }
    }
#undef HPV_SLOT

    // ---- publish this CTA's partial gradient ----
    const float dtot = hpv_block_sum(c, s_red, deps);
    float* gp = a.grad_part + (size_t)c.bid * a.grad_stride;
    for (int i = tid; i < a.theta_pad_n; i += T) gp[i] = s_gw[i];
    if (tid == 0) gp[a.theta_pad_n] = dtot;
}

// K3: grad_pad[i] (+)= sum over CTAs of grad_part[c][i], fixed order.  One CTA of 256 threads = 32 entries x 8
// CTA-groups.  accumulate != 0 adds onto the existing value (several losses into one gradient).
struct HpvGradReduceArgs {
    const float* grad_part;
    int n_parts, stride, n;    // n entries (theta_pad_n + 1)
    float* grad_pad;
    int accumulate;
};

HPV_HD void hpv_gradreduce_body(const HpvCta& c, const HpvGradReduceArgs& a) {
    float* s = reinterpret_cast<float*>(c.smem);     // [8][32]
    const int li = c.tid & 31, grp = c.tid >> 5, ngrp = c.nthreads >> 5;
    const int i = c.bid * 32 + li;
    float acc = 0.0f;
    if (i < a.n)
        for (int p = grp; p < a.n_parts; p += ngrp) acc += a.grad_part[(size_t)p * a.stride + i];
    s[grp * 32 + li] = acc;
    hpv_sync(c);
    if (grp == 0 && i < a.n) {
        float t = 0.0f;
        for (int g2 = 0; g2 < ngrp; ++g2) t += s[g2 * 32 + li];
        a.grad_pad[i] = a.accumulate ? a.grad_pad[i] + t : t;
    }
}
