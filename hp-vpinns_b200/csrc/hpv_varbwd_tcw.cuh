// Reverse sweep through the MLP with EVERY product on the tensor cores (sm_100a) -- hpv_varbwd_tc.cuh plus the
// weight gradients:
//   * forward recompute  z_l = h_{l-1} W_l + b_l  and adjoint propagation  hbar_{l-1} = zbar_l W_l^T :
//     tcgen05.mma kind::tf32, M = 128 points, A operand in TMEM, weights as K-major B tiles (as hpv_varbwd_tc.cuh);
//   * weight gradients  dW_l = sum over channels and points of  h_{l-1}^T zbar_l  (and  db_l = sum zbar_l  through a
//     constant-one row): tcgen05.mma kind::tf32 with M = 64, N = 32, K = 8 POINTS per instruction; both operands are
//     "MN-major" (the point index is the contraction index) and are written by the threads that computed them into
//     shared memory in the TF32 transposed-operand layout of the UMMA descriptor (128-byte swizzle with 32-byte base,
//     cute::UMMA::Layout_MN_SW128_32B_Atom; checked on the hardware by tools/probes/umma_probe.cu test 14);
//     the 3-term split stacks h_hi and h_lo along M (rows 0..31 / 32..63) and runs zbar_hi and zbar_lo as two
//     instructions per K step; the accumulators of all hidden layers stay in TMEM for the whole launch and are read
//     once at the end.  One operand buffer serves the channels in turn.
//   * first and output layer (K = 2, N = 1): per-thread register accumulators over all tiles, reduced over the
//     point rows in a fixed order at the end.
// Deterministic like the other kernels: the tensor-core accumulation order is fixed by the issue order.
#pragma once
#include "hpv_varbwd_tc.cuh"

#define HPV_TCW_OPBYTES 65536                   // operand buffer of one channel: A' 32 KB (64 rows x 128 points), B'hi, B'lo 16 KB each

HPV_HD constexpr int hpv_tcw_tmem_need(int nch, int hp, int nhid) {
    return nch * HPV_TC_NPAD + 2 * nch * (((hp + 1 + 7) / 8) * 8) + HPV_TC_NPAD * (nhid > 1 ? nhid - 1 : 0);
}
HPV_HD constexpr bool hpv_tcw_supported(int nch, int hp, int nhid) { return hp <= 24 && nhid >= 2 && hpv_tcw_tmem_need(nch, hp, nhid) <= 512; }

struct HpvBwdTcwSmem {
    int slots, th, part, red, B, bar, op, total;      // offsets in floats (op: 1 KB aligned at run time, 1 KB of slack included)
    int NS, slot_sz, B_layer, th_n;
};

HPV_HD HpvBwdTcwSmem hpv_bwd_tcw_smem(int dim, int hp, int nch, int nhid) {
    HpvBwdTcwSmem s;
    const int kp = ((hp + 1 + 7) / 8) * 8, sp = hpv_sp(hp);
    int o = 0;
    s.NS = nhid - 2 > 0 ? nhid - 2 : 0;                         // mixed states of hidden layers 1 .. top-1
    s.slot_sz = nch * HPV_TC_MTILE * sp;
    s.slots = o; o += s.NS * s.slot_sz;
    s.th_n = hpv_align4((dim + 1) * hp + hp + 4);
    s.th = o; o += s.th_n;
    s.part = o; o += nch * HPV_TC_MTILE;
    s.red = o; o += HPV_THREADS;
    o = (o + 31) & ~31;
    s.B_layer = 2 * kp * HPV_TC_NPAD;
    s.B = o; o += 2 * (nhid - 1 > 0 ? nhid - 1 : 0) * s.B_layer;
    s.bar = o; o += 2 * (HPV_NFIELDS + 1) + 2;
    s.op = o; o += HPV_TCW_OPBYTES / 4 + 256;
    s.total = o;
    return s;
}

#if defined(__CUDACC__)

// The weight-gradient MMAs of one channel: D_l[64][32] += A'[64][128 points] . (B'hi + B'lo)[32][128 points]^T.
template <uint32_t TB>
__device__ __forceinline__ void hpv_tcw_issue_wgrad(uint32_t opA, uint32_t colW, uint64_t* barW) {
    constexpr uint32_t idesc = hpv_umma_idesc_tf32(64, HPV_TC_NPAD, 1, 1);
    const uint32_t opBhi = opA + 32768u, opBlo = opA + 49152u;
#pragma unroll
    for (int ks = 0; ks < HPV_TC_MTILE / 8; ++ks) {
        const uint64_t ad = hpv_umma_desc(opA + ks * 2048u, 512u, 1024u, 1u);
        hpv_umma_ss(TB + colW, ad, hpv_umma_desc(opBhi + ks * 1024u, 512u, 512u, 1u), idesc, 1u);
        hpv_umma_ss(TB + colW, ad, hpv_umma_desc(opBlo + ks * 1024u, 512u, 512u, 1u), idesc, 1u);
    }
    hpv_umma_commit(barW);
}

template <int DIM, int MX, int MY, int HP, int ACT>
__device__ __forceinline__ void hpv_mlpbwd_tcw_body(const HpvCta& c, const HpvBwdArgs& ba) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int NCH = M::NCH, KP = HpvTcDims<HP>::KP, HPH = HpvTcDims<HP>::HPH, NPR = HPH / 2, SP = HpvSP<HP>::value;
    typedef HpvState<DIM, MX, MY, HPH> State;
    const HpvVarArgs& a = ba.v;
    const int T = c.nthreads, tid = c.tid, nhid = a.nhid, top = nhid - 1;
    const int warp = tid >> 5, lane = tid & 31, sub = warp & 3, half = warp >> 2, u0 = half * HPH;
    const int prow = sub * 32 + lane;
    const HpvBwdTcwSmem L = hpv_bwd_tcw_smem(DIM, HP, NCH, nhid);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_red = sm + L.red;
    float* s_th = sm + L.th;
    float* s_part = sm + L.part;
    uint32_t* s_B = reinterpret_cast<uint32_t*>(sm + L.B);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + L.bar);           // [0..NCH): product channels, [HPV_NFIELDS]: weight gradients
    uint32_t* s_tbase = reinterpret_cast<uint32_t*>(s_bar + HPV_NFIELDS + 1);
    float* const slots = sm + L.slots;
#define HPV_P(l) (slots + (size_t)((l) - 1) * L.slot_sz)
    // operand buffer, 1 KB aligned (the swizzle of the transposed-operand layout is a function of the address bits)
    const uint32_t op_s = (hpv_smem_u32(sm + L.op) + 1023u) & ~1023u;
    unsigned char* const op = reinterpret_cast<unsigned char*>(sm + L.op) + (op_s - hpv_smem_u32(sm + L.op));
    const uint32_t tcols = hpv_tcw_tmem_need(NCH, HP, nhid) <= 256 ? 256u : 512u;

    if (warp == 0) hpv_tmem_alloc(s_tbase, tcols);
    if (tid == 0) {
        for (int i = 0; i <= HPV_NFIELDS; ++i) hpv_mbar_init(&s_bar[i], 1);
        hpv_mbar_init_fence();
    }
    for (int i = tid; i < HPV_TCW_OPBYTES / 16; i += T) reinterpret_cast<uint4*>(op)[i] = make_uint4(0u, 0u, 0u, 0u);
    {
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
            if (which == 0) { if (n < HP) v = k < HP ? W[k * HP + n] : (k == HP ? W[HP * HP + n] : 0.0f); }
            else { if (n < HP && k < HP) v = W[n * HP + k]; }
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
    hpv_fence_proxy_async();
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = *s_tbase;
    if (tcols == 256u ? (tb != 0u && tb != 256u) : (tb != 0u)) { asm volatile("trap;"); }
    const uint32_t tb_lane = tb + ((uint32_t)(sub * 32) << 16);
    constexpr uint32_t colAhi = NCH * HPV_TC_NPAD, colAlo = colAhi + NCH * KP, colW = colAlo + NCH * KP;
    if (half == 0) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
            for (int k = HP; k < KP; k += 2) {
                hpv_tmem_st2(tb_lane + colAhi + ch * KP + k, (ch == 0 && k == HP) ? __float_as_uint(1.0f) : 0u, 0u);
                hpv_tmem_st2(tb_lane + colAlo + ch * KP + k, 0u, 0u);
            }
        }
        // weight-gradient accumulators: zero (they accumulate over every tile of this CTA)
        for (int col = 0; col < top * HPV_TC_NPAD; col += 8) {
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            hpv_tmem_st8(tb_lane + colW + col, z);
        }
        hpv_tmem_wait_st();
    }
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();

    const int Q = a.Q, npts_el = a.rows * Q;
    float deps = 0.0f;
    const long long n_tiles = ((long long)ba.n_points + HPV_TC_MTILE - 1) / HPV_TC_MTILE;
    const int t_begin = (int)(((long long)c.bid * n_tiles) / c.nblocks), t_end = (int)(((long long)(c.bid + 1) * n_tiles) / c.nblocks);
    const float* W1 = s_th;
    const float* b1 = s_th + DIM * HP;
    const float* Wo = s_th + (DIM + 1) * HP;
    uint32_t phase = 0, phaseW = 0;
    bool w_pending = false;                              // weight-gradient MMAs in flight that read the operand buffer

    // register accumulators of the first- and output-layer gradients (this thread's units, summed over its tiles)
    float acc_wo[HPH], acc_w1x[HPH], acc_w1y[HPH], acc_b1[HPH], acc_bo = 0.0f;
#pragma unroll
    for (int j = 0; j < HPH; ++j) { acc_wo[j] = 0.0f; acc_w1x[j] = 0.0f; acc_w1y[j] = 0.0f; acc_b1[j] = 0.0f; }

    // this thread's place in the transposed-operand layout: point k = prow -> k-block kb, row kk of the 4-row atom
    const int kb = prow >> 2, kk = prow & 3;
    unsigned char* const opA_t = op + kb * 1024 + kk * 128;            // A': two row blocks of 512 B per k-block
    unsigned char* const opBh_t = op + 32768 + kb * 512 + kk * 128;
    unsigned char* const opBl_t = op + 49152 + kb * 512 + kk * 128;
    auto sw = [&](int r) -> int { return ((((r & 31) >> 3) ^ kk) << 5) + ((r & 7) << 2); };    // byte offset of row r within its 128-byte line

    auto issue_products = [&](int tile_index) {
        if (warp == 0) {
            hpv_tc_fence_after();
            if (hpv_elect_one()) {
                const uint32_t bhi = hpv_smem_u32(s_B) + (uint32_t)tile_index * (uint32_t)(L.B_layer * 4);
                const uint32_t blo = bhi + (uint32_t)(KP * HPV_TC_NPAD * 4);
                if (tb == 0u) hpv_tc_issue_layer<0, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
                else hpv_tc_issue_layer<256, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
            }
            __syncwarp();
        }
    };
    auto issue_wgrad = [&](int l) {
        if (warp == 0) {
            hpv_tc_fence_after();
            if (hpv_elect_one()) {
                const uint32_t cw = colW + (uint32_t)(l - 1) * HPV_TC_NPAD;
                if (tb == 0u) hpv_tcw_issue_wgrad<0>(op_s, cw, &s_bar[HPV_NFIELDS]);
                else hpv_tcw_issue_wgrad<256>(op_s, cw, &s_bar[HPV_NFIELDS]);
            }
            __syncwarp();
        }
    };
    // operands of the weight-gradient product of one channel: h (left factor: hi rows u, lo rows 32+u, and the bias
    // row HP = 1 for the value channel) and zbar (right factor, hi and lo)
    auto store_operands = [&](const hpv_pair* hp_, const hpv_pair* zp_, int ch) {
#pragma unroll
        for (int m = 0; m < NPR; ++m) {
            const int u = u0 + 2 * m;
            float h0, h1, z0, z1;
            hpv_unpack(hp_[m], h0, h1);
            hpv_unpack(zp_[m], z0, z1);
            uint32_t hh0, hl0, hh1, hl1, zh0, zl0, zh1, zl1;
            hpv_split_trunc(h0, hh0, hl0); hpv_split_trunc(h1, hh1, hl1);
            hpv_split_trunc(z0, zh0, zl0); hpv_split_trunc(z1, zh1, zl1);
            const int o = sw(u);
            *reinterpret_cast<uint2*>(opA_t + o) = make_uint2(hh0, hh1);
            *reinterpret_cast<uint2*>(opA_t + 512 + o) = make_uint2(hl0, hl1);
            *reinterpret_cast<uint2*>(opBh_t + o) = make_uint2(zh0, zh1);
            *reinterpret_cast<uint2*>(opBl_t + o) = make_uint2(zl0, zl1);
        }
        if (half == 1) *reinterpret_cast<uint2*>(opA_t + sw(HP)) = make_uint2(ch == 0 ? __float_as_uint(1.0f) : 0u, 0u);
    };

#pragma unroll 1
    for (int tile = t_begin; tile < t_end; ++tile) {
        const long long gpl = (long long)tile * HPV_TC_MTILE + prow;
        const bool valid = gpl < (long long)ba.n_points;
        const int gp = valid ? (int)gpl : 0;
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
#pragma unroll
            for (int t = 0; t < HPV_MAX_TERMS; ++t)
                if (t < a.n_terms) gbar[t] = ba.Gbar[(size_t)t * ba.n_points + gp];
        }
        float gf[HPV_NFIELDS];
#pragma unroll
        for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
            for (int k = 0; k < HPV_NFIELDS; ++k)
                if (t < a.n_terms) gf[k] = fmaf(gbar[t], fmaf(eps, a.terms[t].a1[k], a.terms[t].a0[k]), gf[k]);
        const float vx = gf[1], vy = gf[2];
        if constexpr (M::DIR) { gf[1] = 1.0f; gf[2] = 0.0f; }

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
        State pre, g, hh;
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
            issue_products(l - 1);
            hpv_tc_load_D<M, HPH>(tb_lane, u0, s_bar, phase, pre);          // pre-activations of layer l
            phase ^= 1;
        }
        hpv_to_mixed<DIM, MX, MY, HPH, ACT>(pre);                           // mixed state of the top layer from here on
        g = pre;
        hpv_activate<DIM, MX, MY, HPH, ACT, true>(g);                       // h_top
        // ---- output layer: Wo/bo gradient into the register accumulators, adjoint of h_top ----
        {
            float go[NCH];
            go[M::C_V] = gf[0];
            if constexpr (M::DX) go[M::C_DX] = gf[1];
            if constexpr (M::DY) go[M::C_DY] = gf[2];
            if constexpr (M::EX) go[M::C_EX] = gf[3];
            if constexpr (M::EY) go[M::C_EY] = gf[4];
            float facc[NCH];
            hpv_each_ch<M>(g, [&](hpv_pair* hp_, int ch) {
                float sacc = 0.0f;
#pragma unroll
                for (int m = 0; m < NPR; ++m) {
                    float h0, h1;
                    hpv_unpack(hp_[m], h0, h1);
                    acc_wo[2 * m] = fmaf(h0, go[ch], acc_wo[2 * m]);
                    acc_wo[2 * m + 1] = fmaf(h1, go[ch], acc_wo[2 * m + 1]);
                    if constexpr (!M::DIR) { sacc = fmaf(h0, Wo[u0 + 2 * m], sacc); sacc = fmaf(h1, Wo[u0 + 2 * m + 1], sacc); }
                }
                facc[ch] = sacc;
            });
            if (half == 0) acc_bo += gf[0];
            if constexpr (!M::DIR) {
                // the fields themselves (d loss / d eps): output sums over the units of both halves
                if (half == 1) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) s_part[ch * HPV_TC_MTILE + prow] = facc[ch];
                }
                __syncthreads();
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
                        for (int k = 0; k < HPV_NFIELDS; ++k) d1 = fmaf(t < a.n_terms ? a.terms[t].a1[k] : 0.0f, f[k], d1);
                        deps = fmaf(gbar[t], d1, deps);
                    }
                }
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
        }

        // ---- hidden layers, top down ----
#pragma unroll 1
        for (int l = top; l >= 1; --l) {
            hpv_activate_bwd<DIM, MX, MY, HPH, ACT, true>(pre, g);           // g := zbar_l, adjoint of the pre-activations of layer l
            hpv_tc_store_A<M, HPH, KP>(tb_lane, u0, g);                      // A operand of the adjoint product
            if (l - 1 >= 1) hpv_tc_load_half<M, HPH, SP>(HPV_P(l - 1), prow, u0, pre);       // mixed state of layer l-1
            else { layer1(pre); hpv_to_mixed<DIM, MX, MY, HPH, ACT>(pre); }
            hh = pre;
            hpv_activate<DIM, MX, MY, HPH, ACT, true>(hh);                   // h_{l-1}: left factor of the W_l gradient
            // channel 0 of the weight-gradient product goes out together with the adjoint product
            if (w_pending) { hpv_mbar_wait(&s_bar[HPV_NFIELDS], phaseW); phaseW ^= 1; w_pending = false; }
            store_operands(hh.v.p, g.v.p, 0);
            hpv_fence_proxy_async();
            hpv_tmem_wait_st();
            hpv_tc_fence_before();
            __syncthreads();
            issue_products(top + (l - 1));                                   // hbar_{l-1} = zbar_l . W_l^T
            issue_wgrad(l);
            w_pending = true;
            // remaining channels, one at a time through the same operand buffer
            auto more = [&](const hpv_pair* hp_, const hpv_pair* zp_, int ch) {
                hpv_mbar_wait(&s_bar[HPV_NFIELDS], phaseW); phaseW ^= 1;
                hpv_tc_fence_after();
                store_operands(hp_, zp_, ch);
                hpv_fence_proxy_async();
                __syncthreads();
                issue_wgrad(l);
            };
            if constexpr (M::DX) more(hh.dx.p, g.dx.p, M::C_DX);
            if constexpr (M::DY) more(hh.dy.p, g.dy.p, M::C_DY);
            if constexpr (M::EX) more(hh.ex.p, g.ex.p, M::C_EX);
            if constexpr (M::EY) more(hh.ey.p, g.ey.p, M::C_EY);
            hpv_tc_load_D<M, HPH>(tb_lane, u0, s_bar, phase, g);             // adjoint of h_{l-1}
            phase ^= 1;
        }

        // ---- first layer: W1/b1 gradient into the register accumulators ----
        hpv_activate_bwd<DIM, MX, MY, HPH, ACT, true>(pre, g);               // g := zbar_0
#pragma unroll
        for (int m = 0; m < NPR; ++m) {
            float zv0, zv1;
            hpv_unpack(g.v.p[m], zv0, zv1);
            acc_b1[2 * m] += zv0; acc_b1[2 * m + 1] += zv1;
            acc_w1x[2 * m] = fmaf(x, zv0, acc_w1x[2 * m]); acc_w1x[2 * m + 1] = fmaf(x, zv1, acc_w1x[2 * m + 1]);
            if constexpr (DIM == 2) { acc_w1y[2 * m] = fmaf(y, zv0, acc_w1y[2 * m]); acc_w1y[2 * m + 1] = fmaf(y, zv1, acc_w1y[2 * m + 1]); }
            if constexpr (M::DIR) {                                           // tangent seed v . W1: left factor v
                float zd0, zd1;
                hpv_unpack(g.dx.p[m], zd0, zd1);
                acc_w1x[2 * m] = fmaf(vx, zd0, acc_w1x[2 * m]); acc_w1x[2 * m + 1] = fmaf(vx, zd1, acc_w1x[2 * m + 1]);
                acc_w1y[2 * m] = fmaf(vy, zd0, acc_w1y[2 * m]); acc_w1y[2 * m + 1] = fmaf(vy, zd1, acc_w1y[2 * m + 1]);
            } else {
                if constexpr (M::DX) { float d0, d1; hpv_unpack(g.dx.p[m], d0, d1); acc_w1x[2 * m] += d0; acc_w1x[2 * m + 1] += d1; }
                if constexpr (M::DY) { float d0, d1; hpv_unpack(g.dy.p[m], d0, d1); acc_w1y[2 * m] += d0; acc_w1y[2 * m + 1] += d1; }
            }
        }
    }
#undef HPV_P
    if (w_pending) { hpv_mbar_wait(&s_bar[HPV_NFIELDS], phaseW); phaseW ^= 1; w_pending = false; }
    hpv_tc_fence_after();
    hpv_pdl_trigger();

    // ---- publish this CTA's partial gradient (padded layout) ----
    float* gpart = a.grad_part + (size_t)c.bid * a.grad_stride;
    const float dtot = hpv_block_sum(c, s_red, deps);
    for (int ip = tid; ip < a.theta_pad_n; ip += T) gpart[ip] = 0.0f;
    // (a) first and output layer: the threads' register accumulators, summed over the 128 point rows in a fixed order
    float* st = reinterpret_cast<float*>(op);                               // staging in the (now idle) operand buffer
    constexpr int NA = 4 * HPH + 1;
    __syncthreads();
    {
        float* row = st + (size_t)tid * NA;
#pragma unroll
        for (int j = 0; j < HPH; ++j) { row[j] = acc_wo[j]; row[HPH + j] = acc_w1x[j]; row[2 * HPH + j] = acc_w1y[j]; row[3 * HPH + j] = acc_b1[j]; }
        row[4 * HPH] = acc_bo;
    }
    __syncthreads();
    for (int o = tid; o < 2 * NA; o += T) {
        const int h2 = o / NA, j = o - h2 * NA;                            // half, accumulator index
        float sum = 0.0f;
        for (int s2 = 0; s2 < 4; ++s2)
            for (int ln = 0; ln < 32; ++ln) sum += st[(size_t)((h2 * 4 + s2) * 32 + ln) * NA + j];
        const int u = h2 * HPH + (j % HPH), kind = j / HPH;
        if (j == 4 * HPH) { if (h2 == 0) gpart[a.off_wo + HP] = sum; }
        else if (kind == 0) gpart[a.off_wo + u] = sum;
        else if (kind == 1) gpart[u] = sum;
        else if (kind == 2) { if (DIM == 2) gpart[HP + u] = sum; }
        else gpart[DIM * HP + u] = sum;
    }
    __syncthreads();
    // (b) hidden layers: the TMEM accumulators  D_l[r][j], rows r < 32 from h_hi, rows 32 + r from h_lo
    for (int l = 1; l <= top; ++l) {
        if (warp < 4) {
            // M = 64 accumulator: row r lives in lane (r % 16) + 32 (r / 16): lanes 0..15 of sub-partition `warp` hold rows 16 warp ..
            float v[HPV_TC_NPAD];
#pragma unroll
            for (int c8 = 0; c8 < HPV_TC_NPAD; c8 += 8) {
                uint32_t r8[8];
                hpv_tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + colW + (uint32_t)(l - 1) * HPV_TC_NPAD + c8, r8);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[c8 + j] = __uint_as_float(r8[j]);
            }
            hpv_tmem_wait_ld();
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < HPV_TC_NPAD; ++j) st[(warp * 16 + lane) * HPV_TC_NPAD + j] = v[j];
            }
        }
        __syncthreads();
        float* gW = gpart + hpv_off_wl(DIM, HP, l);
        for (int o = tid; o < (HP + 1) * HP; o += T) {
            const int r = o / HP, j = o - r * HP;                           // r = HP: the bias row (follows W_l in the padded layout)
            gW[o] = st[r * HPV_TC_NPAD + j] + st[(32 + r) * HPV_TC_NPAD + j];
        }
        __syncthreads();
    }
    if (tid == 0) gpart[a.theta_pad_n] = dtot;
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, tcols);
}

#endif  // __CUDACC__
