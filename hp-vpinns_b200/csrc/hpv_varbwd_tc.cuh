// Reverse sweep through the MLP, tensor-core form (sm_100a): the same algorithm as hpv_mlpbwd_body (hpv_varbwd.cuh) --
// recompute the forward-mode MLP at the quadrature points, run the hand-derived reverse sweep through it, accumulate
// the weight gradients over the points (what AdamOptimizer.minimize differentiates, P2D:131-132) -- with every
// hidden-layer product on the tensor cores: the forward recompute  z_l = h_{l-1} W_l + b_l  and the adjoint
// propagation  hbar_{l-1} = zbar_l W_l^T  are tcgen05.mma kind::tf32 with the activations / adjoints as the A operand
// in tensor memory (3-term TF32 split, as in hpv_varfwd_tc.cuh) and the weights -- and, for the adjoint product, the
// same matrix read the other way round -- as K-major B tiles in shared memory.  The weight-gradient products stay on
// the FMA pipe (hpv_wgrad_warp) and run while the adjoint MMAs of the same layer execute.
// Thread layout as in the tensor-core forward kernel: CTA = 256 threads, 128 points per tile, warp w on TMEM
// sub-partition w % 4 and the units [HP/2 (w/4), HP/2 (w/4) + HP/2); the two warps that share 32 points split the
// weight-gradient GEMMs by points (16 each) into their own accumulators.
#pragma once
#include "hpv_varbwd.cuh"
#include "hpv_varfwd_tc.cuh"

struct HpvBwdTcSmem {
    int cst, slots, go, in0, gw, red, th, part, B, bar, total;      // offsets in floats
    int NS, slot_sz, gwn, B_layer, th_n;
};

HPV_HD HpvBwdTcSmem hpv_bwd_tc_smem(int dim, int hp, int nch, int nch1, int nhid) {
    HpvBwdTcSmem s;
    const int kp = ((hp + 1 + 7) / 8) * 8, sp = hpv_sp(hp);
    int o = 0;
    s.cst = o; o += 8;
    s.NS = nhid - 1 > 2 ? nhid - 1 : 2;
    s.slot_sz = nch * HPV_TC_MTILE * sp;
    s.slots = o; o += s.NS * s.slot_sz;
    s.go = o; o += nch * HPV_TC_MTILE;
    s.in0 = o; o += nch1 * HPV_TC_MTILE * 4;
    s.gwn = hpv_gw_n(dim, hp, nhid);
    s.gw = o; o += (HPV_THREADS / 32) * s.gwn;
    s.red = o; o += HPV_THREADS;
    s.th_n = hpv_align4((dim + 1) * hp + hp + 4);
    s.th = o; o += s.th_n;
    s.part = o; o += nch * HPV_TC_MTILE;
    o = (o + 31) & ~31;
    s.B_layer = 2 * kp * HPV_TC_NPAD;                          // hi tile, lo tile
    s.B = o; o += 2 * (nhid - 1 > 0 ? nhid - 1 : 0) * s.B_layer;   // forward tiles, then adjoint tiles
    s.bar = o; o += 2 * HPV_NFIELDS + 2;
    s.total = o;
    return s;
}

#if defined(__CUDACC__)

// This thread's units of one channel block [ch][128 rows][SP] <-> registers (64-bit accesses of the packed pairs).
// 128-bit accesses wherever the 16-byte alignment allows (rows are SP floats apart, SP chosen so that the 128-bit
// accesses of a quarter warp fall into distinct banks; plain 64-bit accesses of the pairs measured 4-way conflicts):
// HPH = 10: units 0-9 = quad, quad, pair; units 10-19 = pair, quad, quad.
template <class M, int HPH, int SP, class S>
__device__ __forceinline__ void hpv_tc_store_half(float* slot, int prow, int u0, const S& s) {
    constexpr int NPRS = HPH / 2;
    hpv_each_ch<M>(s, [&](const hpv_pair* a, int ch) {
        float* row = slot + ((size_t)ch * HPV_TC_MTILE + prow) * SP + u0;
        if constexpr (NPRS % 2 == 0) {
#pragma unroll
            for (int m = 0; m < NPRS; m += 2) hpv_st_pairs(row + 2 * m, a[m], a[m + 1]);
        } else if ((u0 & 3) == 0) {                         // quads first, one trailing pair
#pragma unroll
            for (int m = 0; m + 1 < NPRS; m += 2) hpv_st_pairs(row + 2 * m, a[m], a[m + 1]);
            *reinterpret_cast<hpv_pair*>(row + 2 * (NPRS - 1)) = a[NPRS - 1];
        } else {                                            // one leading pair, then quads
            *reinterpret_cast<hpv_pair*>(row) = a[0];
#pragma unroll
            for (int m = 1; m + 1 < NPRS; m += 2) hpv_st_pairs(row + 2 * m, a[m], a[m + 1]);
        }
    });
}
template <class M, int HPH, int SP, class S>
__device__ __forceinline__ void hpv_tc_load_half(const float* slot, int prow, int u0, S& s) {
    constexpr int NPRS = HPH / 2;
    hpv_each_ch<M>(s, [&](hpv_pair* a, int ch) {
        const float* row = slot + ((size_t)ch * HPV_TC_MTILE + prow) * SP + u0;
        if constexpr (NPRS % 2 == 0) {
#pragma unroll
            for (int m = 0; m < NPRS; m += 2) hpv_ld_pairs(row + 2 * m, a[m], a[m + 1]);
        } else if ((u0 & 3) == 0) {
#pragma unroll
            for (int m = 0; m + 1 < NPRS; m += 2) hpv_ld_pairs(row + 2 * m, a[m], a[m + 1]);
            a[NPRS - 1] = *reinterpret_cast<const hpv_pair*>(row + 2 * (NPRS - 1));
        } else {
            a[0] = *reinterpret_cast<const hpv_pair*>(row);
#pragma unroll
            for (int m = 1; m + 1 < NPRS; m += 2) hpv_ld_pairs(row + 2 * m, a[m], a[m + 1]);
        }
    });
}
// Split this thread's units of every channel and store them as the A operand (hi, lo) of the next product.
template <class M, int HPH, int KP, class S>
__device__ __forceinline__ void hpv_tc_store_A(uint32_t tb_lane, int u0, const S& s) {
    constexpr uint32_t colAhi = M::NCH * HPV_TC_NPAD, colAlo = colAhi + M::NCH * KP;
    hpv_each_ch<M>(s, [&](const hpv_pair* a, int ch) {
        hpv_tc_split_store<HPH>(tb_lane + colAhi + ch * KP + u0, tb_lane + colAlo + ch * KP + u0, a);
    });
}
// The accumulators of a product, channel by channel as they complete -> this thread's units.
template <class M, int HPH, class S>
__device__ __forceinline__ void hpv_tc_load_D(uint32_t tb_lane, int u0, uint64_t* bar, uint32_t phase, S& s) {
    hpv_each_ch<M>(s, [&](hpv_pair* zp, int ch) {
#if !defined(HPV_EXP_NO_MMA)      // timing experiment only (tools/gpu_r2k.sh): no products issued, nothing to wait for
        hpv_mbar_wait(&bar[ch], phase);
#endif
        hpv_tc_fence_after();
        float v[HPH];
        hpv_tmem_ld_n<HPH>(tb_lane + ch * HPV_TC_NPAD + u0, v);
        hpv_tmem_wait_ld();
#pragma unroll
        for (int m = 0; m < HPH / 2; ++m) zp[m] = hpv_pack(v[2 * m], v[2 * m + 1]);
    });
}

template <int DIM, int MX, int MY, int HP, int ACT>
__device__ __forceinline__ void hpv_mlpbwd_tc_body(const HpvCta& c, const HpvBwdArgs& ba) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int NCH = M::NCH, KP = HpvTcDims<HP>::KP, HPH = HpvTcDims<HP>::HPH, NPR = HPH / 2, SP = HpvSP<HP>::value;
    constexpr int NCH1 = 1 + (M::DX ? 1 : 0) + (M::DY ? 1 : 0);
    constexpr uint32_t TCOLS = (NCH * HPV_TC_NPAD + 2 * NCH * KP) <= 256 ? 256u : 512u;
    typedef HpvState<DIM, MX, MY, HPH> State;
    const HpvVarArgs& a = ba.v;
    const int T = c.nthreads, tid = c.tid, nhid = a.nhid, top = nhid - 1;
    const int warp = tid >> 5, lane = tid & 31, sub = warp & 3, half = warp >> 2, u0 = half * HPH, nwarps = T >> 5;
    const int prow = sub * 32 + lane;
    const HpvBwdTcSmem L = hpv_bwd_tc_smem(DIM, HP, NCH, NCH1, nhid);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_cst = sm + L.cst;
    float* s_gw = sm + L.gw + (size_t)warp * L.gwn;
    float* s_go = sm + L.go;
    float* s_in0 = sm + L.in0;
    float* s_red = sm + L.red;
    float* s_th = sm + L.th;
    float* s_part = sm + L.part;
    uint32_t* s_B = reinterpret_cast<uint32_t*>(sm + L.B);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + L.bar);
    uint32_t* s_tbase = reinterpret_cast<uint32_t*>(s_bar + HPV_NFIELDS);
    float* const slots = sm + L.slots;
    float* const X = slots + (size_t)(top >= 1 ? top - 1 : 0) * L.slot_sz;       // running slot (see hpv_mlpbwd_body)
    float* const H0 = slots + (size_t)(top >= 2 ? 0 : 1) * L.slot_sz;
#define HPV_P(l) (slots + (size_t)((l) - 1) * L.slot_sz)

    if (warp == 0) hpv_tmem_alloc(s_tbase, TCOLS);
    if (tid == 0) {
        for (int i = 0; i < HPV_NFIELDS; ++i) hpv_mbar_init(&s_bar[i], 1);
        hpv_mbar_init_fence();
    }
    for (int i = tid; i < (T >> 5) * L.gwn; i += T) sm[L.gw + i] = 0.0f;
    if (tid < 8) s_cst[tid] = (tid == 0) ? 1.0f : 0.0f;
    {
        // parameters from global memory: W1, b1 | Wo, bo into shared memory; the hidden matrices as B tiles
        const float* tg = a.theta_pad;
        const int n1 = (DIM + 1) * HP;
        for (int i = tid; i < n1; i += T) s_th[i] = tg[i];
        for (int i = tid; i < HP + 4; i += T) s_th[n1 + i] = tg[a.off_wo + i];
        const int per_layer = KP * HPV_TC_NPAD;
        for (int i = tid; i < 2 * top * per_layer; i += T) {
            const int which = i / (top * per_layer), r0 = i - which * top * per_layer;
            const int l = r0 / per_layer, r = r0 - l * per_layer, n = r / KP, k = r - n * KP;
            const float* W = tg + hpv_off_wl(DIM, HP, l + 1);
            float v = 0.0f;
            if (which == 0) {          // forward: B[n = out][k = in] = W[in][out], bias as input row HP
                if (n < HP) v = k < HP ? W[k * HP + n] : (k == HP ? W[HP * HP + n] : 0.0f);
            } else {                   // adjoint: B[n = in][k = out] = W[in][out]
                if (n < HP && k < HP) v = W[n * HP + k];
            }
            uint32_t hi, lo;
            hpv_split_trunc(v, hi, lo);
            const int w = (k >> 2) * (HPV_TC_NPAD * 4) + n * 4 + (k & 3);
            uint32_t* dst = s_B + (size_t)(which * top + l) * L.B_layer;
            dst[w] = hi;
            dst[per_layer + w] = lo;
        }
    }
    hpv_pdl_wait();                                      // Gbar (or the point adjoints) of the previous kernel
    const float eps = a.eps[0];
    float coef[HPV_MAX_TERMS][HPV_NFIELDS], coef1[HPV_MAX_TERMS][HPV_NFIELDS];
#pragma unroll
    for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
        for (int f = 0; f < HPV_NFIELDS; ++f) {
            coef[t][f] = (t < a.n_terms) ? fmaf(eps, a.terms[t].a1[f], a.terms[t].a0[f]) : 0.0f;
            coef1[t][f] = (t < a.n_terms) ? a.terms[t].a1[f] : 0.0f;
        }
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = *s_tbase;
    if (TCOLS == 256u ? (tb != 0u && tb != 256u) : (tb != 0u)) { asm volatile("trap;"); }
    const uint32_t tb_lane = tb + ((uint32_t)(sub * 32) << 16);
    constexpr uint32_t colAhi = NCH * HPV_TC_NPAD, colAlo = colAhi + NCH * KP;
    if (half == 0) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
            for (int k = HP; k < KP; k += 2) {
                hpv_tmem_st2(tb_lane + colAhi + ch * KP + k, (ch == 0 && k == HP) ? __float_as_uint(1.0f) : 0u, 0u);
                hpv_tmem_st2(tb_lane + colAlo + ch * KP + k, 0u, 0u);
            }
        }
        hpv_tmem_wait_st();
    }

    const int Q = a.Q, npts_el = a.rows * Q;
    float deps = 0.0f;
    const long long n_tiles = ((long long)ba.n_points + HPV_TC_MTILE - 1) / HPV_TC_MTILE;
    const int t_begin = (int)(((long long)c.bid * n_tiles) / c.nblocks), t_end = (int)(((long long)(c.bid + 1) * n_tiles) / c.nblocks);
    const float* W1 = s_th;
    const float* b1 = s_th + DIM * HP;
    const float* Wo = s_th + (DIM + 1) * HP;
    const int g_wo = hpv_gw_wo(DIM, HP, nhid);
    const int rows16 = sub * 32 + 16 * half;             // the 16 points this warp takes in the weight-gradient GEMMs
    uint32_t phase = 0;

    auto issue = [&](int tile_index) {
#if defined(HPV_EXP_NO_MMA)
        if (tile_index >= 0) return;
#endif
        if (warp == 0) {
            hpv_tc_fence_after();
            if (hpv_elect_one()) {
                const uint32_t bhi = hpv_smem_u32(s_B) + (uint32_t)tile_index * (uint32_t)(L.B_layer * 4);
                const uint32_t blo = bhi + (uint32_t)(KP * HPV_TC_NPAD * 4);
                if (TCOLS == 512u || tb == 0u) hpv_tc_issue_layer<0, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
                else hpv_tc_issue_layer<256, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
            }
            __syncwarp();
        }
    };

    // coordinates and field adjoints of this thread's point of a tile; fetched one tile ahead (the loads would
    // otherwise sit exposed at the top of every tile: all warps of the CTA start a tile in lockstep)
    auto load_inputs = [&](int tile, float& x, float& y, float (&gb)[HPV_MAX_TERMS]) {
        const long long gpl = (long long)tile * HPV_TC_MTILE + prow;
        x = 0.0f; y = 0.0f;
#pragma unroll
        for (int t = 0; t < HPV_MAX_TERMS; ++t) gb[t] = 0.0f;
        if (tile < t_end && gpl < (long long)ba.n_points) {
            const int gp = (int)gpl;
            if (ba.pts) {
                x = ba.pts[(size_t)gp * DIM];
                if (DIM == 2) y = ba.pts[(size_t)gp * DIM + 1];
            } else {
                const int e = gp / npts_el, p = gp - e * npts_el;
                const int j = p / Q, i = p - j * Q;
                x = fmaf(a.el_geom[4 * e + 1], a.xi1[i], a.el_geom[4 * e + 0]);
                if (DIM == 2) y = fmaf(a.el_geom[4 * e + 3], a.xi1[j], a.el_geom[4 * e + 2]);
            }
#pragma unroll
            for (int t = 0; t < HPV_MAX_TERMS; ++t)
                if (t < a.n_terms) gb[t] = ba.Gbar[(size_t)t * ba.n_points + gp];
        }
    };
    float nx, ny, ngbar[HPV_MAX_TERMS];
    load_inputs(t_begin, nx, ny, ngbar);

#pragma unroll 1
    for (int tile = t_begin; tile < t_end; ++tile) {
        const float x = nx, y = ny;
        float gbar[HPV_MAX_TERMS];
#pragma unroll
        for (int t = 0; t < HPV_MAX_TERMS; ++t) gbar[t] = ngbar[t];
        load_inputs(tile + 1, nx, ny, ngbar);
        float gf[HPV_NFIELDS];
#pragma unroll
        for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
            for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = fmaf(gbar[t], coef[t][k], gf[k]);
        const float vx = gf[1], vy = gf[2];
        if constexpr (M::DIR) { gf[1] = 1.0f; gf[2] = 0.0f; }

        // first layer, this thread's units (directional mode: tangent seed v . W1)
        auto layer1 = [&](State& z) {
            const hpv_pair xx = hpv_dup(x), yy = hpv_dup(y);
#pragma unroll
            for (int m = 0; m < NPR; ++m) {
                const int u = u0 + 2 * m;
                const hpv_pair wx = hpv_pack(W1[u], W1[u + 1]);
                hpv_pair zz = hpv_fma2r(xx, wx, hpv_pack(b1[u], b1[u + 1]));
                if constexpr (DIM == 2) {
                    const hpv_pair wy = hpv_pack(W1[HP + u], W1[HP + u + 1]);
                    zz = hpv_fma2r(yy, wy, zz);
                    if constexpr (M::DIR) z.dx.p[m] = hpv_fma2r(hpv_dup(vy), wy, hpv_mul2(hpv_dup(vx), wx));
                    if constexpr (M::DY) z.dy.p[m] = wy;
                }
                z.v.p[m] = zz;
                if constexpr (M::DX && !M::DIR) z.dx.p[m] = wx;
                if constexpr (M::EX) z.ex.p[m] = hpv_dup(0.0f);
                if constexpr (M::EY) z.ey.p[m] = hpv_dup(0.0f);
            }
        };

        // ---- forward recompute; the mixed states of hidden layers 1..top-1 stay in their slots ----
        State pre, g;
        layer1(pre);
#pragma unroll 1
        for (int l = 1; l <= top; ++l) {
            hpv_to_mixed<DIM, MX, MY, HPH, ACT>(pre);                       // mixed state of layer l-1
            if (l >= 2) hpv_tc_store_half<M, HPH, SP>(HPV_P(l - 1), prow, u0, pre);
            g = pre;
            hpv_activate<DIM, MX, MY, HPH, ACT, true>(g);                   // h_{l-1}
            hpv_tc_store_A<M, HPH, KP>(tb_lane, u0, g);
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            issue(l - 1);
            hpv_tc_load_D<M, HPH>(tb_lane, u0, s_bar, phase, pre);          // pre-activations of layer l
            phase ^= 1;
        }
        hpv_to_mixed<DIM, MX, MY, HPH, ACT>(pre);                           // mixed state of the top layer from here on
        g = pre;
        hpv_activate<DIM, MX, MY, HPH, ACT, true>(g);                       // h_top
        hpv_tc_store_half<M, HPH, SP>(X, prow, u0, g);                      // left factor of the Wo gradient
        float facc[NCH];
        if constexpr (!M::DIR) {
            // the fields themselves (d loss / d eps): partial output sums over this thread's units
            hpv_each_ch<M>(g, [&](hpv_pair* hp_, int ch) {
                float sacc = 0.0f;
#pragma unroll
                for (int m = 0; m < NPR; ++m) {
                    float h0, h1;
                    hpv_unpack(hp_[m], h0, h1);
                    sacc = fmaf(h0, Wo[u0 + 2 * m], sacc);
                    sacc = fmaf(h1, Wo[u0 + 2 * m + 1], sacc);
                }
                facc[ch] = sacc;
            });
            if (half == 1) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) s_part[ch * HPV_TC_MTILE + prow] = facc[ch];
            }
        }
        // ---- output layer: adjoint of h_top, Wo/bo gradient ----
        if (half == 0) {
            s_go[M::C_V * HPV_TC_MTILE + prow] = gf[0];
            if constexpr (M::DX) s_go[M::C_DX * HPV_TC_MTILE + prow] = gf[1];
            if constexpr (M::DY) s_go[M::C_DY * HPV_TC_MTILE + prow] = gf[2];
            if constexpr (M::EX) s_go[M::C_EX * HPV_TC_MTILE + prow] = gf[3];
            if constexpr (M::EY) s_go[M::C_EY * HPV_TC_MTILE + prow] = gf[4];
        }
#pragma unroll
        for (int m = 0; m < NPR; ++m) {
            const hpv_pair w = hpv_pack(Wo[u0 + 2 * m], Wo[u0 + 2 * m + 1]);
            g.v.p[m] = hpv_mul2(hpv_dup(gf[0]), w);
            if constexpr (M::DX) g.dx.p[m] = hpv_mul2(hpv_dup(gf[1]), w);
            if constexpr (M::DY) g.dy.p[m] = hpv_mul2(hpv_dup(gf[2]), w);
            if constexpr (M::EX) g.ex.p[m] = hpv_mul2(hpv_dup(gf[3]), w);
            if constexpr (M::EY) g.ey.p[m] = hpv_mul2(hpv_dup(gf[4]), w);
        }
        __syncthreads();
        if constexpr (!M::DIR) {
            if (half == 0) {
                float f[HPV_NFIELDS];
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) facc[ch] += s_part[ch * HPV_TC_MTILE + prow];
                f[0] = facc[M::C_V] + Wo[HP];
                f[1] = M::DX ? facc[M::DX ? M::C_DX : 0] : 0.0f;
                f[2] = M::DY ? facc[M::DY ? M::C_DY : 0] : 0.0f;
                f[3] = M::EX ? facc[M::EX ? M::C_EX : 0] : 0.0f;
                f[4] = M::EY ? facc[M::EY ? M::C_EY : 0] : 0.0f;
#pragma unroll
                for (int t = 0; t < HPV_MAX_TERMS; ++t) {
                    float d1 = 0.0f;
#pragma unroll
                    for (int k = 0; k < HPV_NFIELDS; ++k) d1 = fmaf(coef1[t][k], f[k], d1);
                    deps = fmaf(gbar[t], d1, deps);
                }
            }
        }
        hpv_wgrad_warp<SP, 1, NCH, HP, 1, 1, true, 1, HP, DIM, 16, HPV_TC_MTILE>(c, X + (size_t)rows16 * SP, s_go + rows16, s_gw + g_wo, nullptr, s_cst);

        // ---- hidden layers, top down ----
#pragma unroll 1
        for (int l = top; l >= 1; --l) {
            __syncthreads();                                                 // everybody is done reading X
            hpv_activate_bwd<DIM, MX, MY, HPH, ACT, true>(pre, g);           // g := adjoint of the pre-activations of layer l
            hpv_tc_store_half<M, HPH, SP>(X, prow, u0, g);                   // right factor of the W_l gradient
            hpv_tc_store_A<M, HPH, KP>(tb_lane, u0, g);                      // A operand of the adjoint product
            float* INl = (l - 1 >= 1) ? HPV_P(l - 1) : H0;
            if (l - 1 >= 1) hpv_tc_load_half<M, HPH, SP>(INl, prow, u0, pre);         // mixed state of layer l-1
            else { layer1(pre); hpv_to_mixed<DIM, MX, MY, HPH, ACT>(pre); }
            g = pre;
            hpv_activate<DIM, MX, MY, HPH, ACT, true>(g);
            hpv_tc_store_half<M, HPH, SP>(INl, prow, u0, g);                 // h_{l-1}: left factor of the W_l gradient
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            issue(top + (l - 1));                                            // hbar_{l-1} = zbar_l . W_l^T on the tensor cores ...
            float* gW = s_gw + hpv_gw_wl(DIM, HP, l);                        // ... while the FMA pipe forms the W_l gradient
            hpv_wgrad_warp<SP, SP, NCH, HP, HP / 4, 4, true, 0, HP, DIM, 16, HPV_TC_MTILE>(c, INl + (size_t)rows16 * SP, X + (size_t)rows16 * SP, gW,
                                                                                          gW + HP * HP, s_cst);
            hpv_tc_load_D<M, HPH>(tb_lane, u0, s_bar, phase, g);
            phase ^= 1;
        }

        // ---- first layer ----
        __syncthreads();
        hpv_activate_bwd<DIM, MX, MY, HPH, ACT, true>(pre, g);
        hpv_tc_store_half<M, HPH, SP>(X, prow, u0, g);
        if (half == 0) {
            HpvF4 o;
            o.x = x; o.y = y; o.z = 1.0f; o.w = 0.0f; hpv_st4(s_in0 + (0 * HPV_TC_MTILE + prow) * 4, o);
            if constexpr (M::DIR) {
                o.x = vx; o.y = vy; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (1 * HPV_TC_MTILE + prow) * 4, o);
            } else {
                if constexpr (M::DX) { o.x = 1.0f; o.y = 0.0f; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (M::C_DX * HPV_TC_MTILE + prow) * 4, o); }
                if constexpr (M::DY) { o.x = 0.0f; o.y = 1.0f; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (M::C_DY * HPV_TC_MTILE + prow) * 4, o); }
            }
        }
        __syncthreads();
        hpv_wgrad_warp<4, SP, NCH1, 4, HP / 4, 4, false, 2, HP, DIM, 16, HPV_TC_MTILE>(c, s_in0 + (size_t)rows16 * 4, X + (size_t)rows16 * SP, s_gw,
                                                                                     s_gw + DIM * HP, s_cst);
        __syncthreads();                                                     // X, s_in0, s_go are rewritten by the next tile
    }
#undef HPV_P

    hpv_pdl_trigger();
    // ---- publish this CTA's partial gradient: the warps' accumulators summed in a fixed order, padded layout ----
    const float dtot = hpv_block_sum(c, s_red, deps);
    float* gpart = a.grad_part + (size_t)c.bid * a.grad_stride;
    const float* gw0 = sm + L.gw;
    for (int ip = tid; ip < a.theta_pad_n; ip += T) {
        const int ic = hpv_gw_of_padded(DIM, HP, nhid, ip);
        float sum = 0.0f;
        if (ic >= 0)
            for (int w = 0; w < nwarps; ++w) sum += gw0[(size_t)w * L.gwn + ic];
        gpart[ip] = sum;
    }
    if (tid == 0) gpart[a.theta_pad_n] = dtot;
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, TCOLS);
}

#endif  // __CUDACC__
