// Backward of the variational loss (what tf.train.AdamOptimizer.minimize differentiates, P2D:131-132,
// P1D:102-104, ADI:191-193), as two kernels:
//   K1  hpv_adjproj_body : adjoint of the projection.  Rbar = dloss/dU = 2 Res / (Ntx Nty) per element, then
//       Gbar_t[j][i] = c_t * sum_k sum_r L_t[j][k] Rbar[k][r] R_t[i][r]  -> adjoint of the point field of term t.
//   K2  hpv_mlpbwd_body  : per quadrature point, recompute the forward-mode MLP, run the reverse sweep through
//       it, and accumulate the weight gradients as small GEMMs over the points of a tile (deterministic).
// A third, tiny kernel (hpv_gradreduce_body) sums the per-CTA partial gradients in a fixed order.
#pragma once
#include "hpv_varfwd.cuh"
#include "hpv_slot.cuh"

#define HPV_ADJ_RS 16          // rows of an element handled by one CTA of K1
#define HPV_ADJ_RSP 20         // padded row count (stride of the transposed V buffer, conflict-free for 128-bit loads)

struct HpvAdjSmem { int tab[HPV_NTAB], Rbar, V, total; };

HPV_HD HpvAdjSmem hpv_adj_smem(const HpvVarArgs& a) {
    HpvAdjSmem s;
    int o = 0;
    int m = hpv_tab_mask(a);
    for (int t = 0; t < HPV_NTAB; ++t) {
        s.tab[t] = -1;
        if (m & (1 << t)) { s.tab[t] = o; o += HPV_NP * a.QP + 4; }
    }
    s.Rbar = o; o += HPV_NP * HPV_NP;
    s.V = o; o += a.n_terms * HPV_NP * HPV_ADJ_RSP;
    s.total = o;
    return s;
}

// Gbar layout: [term][element][rows*Q]
struct HpvAdjArgs {
    HpvVarArgs v;
    float* Gbar;
    int slabs_per_el;
};

// One CTA per (element, slab of HPV_ADJ_RS rows).  Both contractions are register-tiled 4x4 with 128-bit
// shared-memory loads along the contiguous index of the natural-layout tables.
HPV_HD void hpv_adjproj_body(const HpvCta& c, const HpvAdjArgs& aa) {
    const HpvVarArgs& a = aa.v;
    const HpvAdjSmem L = hpv_adj_smem(a);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_R = sm + L.Rbar;
    float* s_V = sm + L.V;
    const int T = c.nthreads, tid = c.tid, Q = a.Q, QP = a.QP;
    const int e = c.bid / aa.slabs_per_el, slab = c.bid - e * aa.slabs_per_el;
    const int j0 = slab * HPV_ADJ_RS;
    int nrows = a.rows - j0;
    if (nrows > HPV_ADJ_RS) nrows = HPV_ADJ_RS;
    const int npts_el = a.rows * Q;

    for (int t = 0; t < HPV_NTAB; ++t) {                 // the tables do not depend on the forward kernel: staged
        if (L.tab[t] < 0) continue;                      // while its last CTAs are still running
        const float* src = a.tabN[t];
        float* dst = sm + L.tab[t];
        for (int i = tid * 4; i < HPV_NP * QP + 4; i += T * 4) hpv_st4(dst + i, hpv_ld4(src + i));
    }
    hpv_pdl_wait();                                      // Res of the forward kernel from here on
    const int ntx_e = a.el_ntest[2 * e + 0], nty_e = a.el_ntest[2 * e + 1];
    const float rs = a.loss_scale * 2.0f / (float)(ntx_e * nty_e);
    constexpr int RPT = HPV_NP * HPV_NP / HPV_THREADS;       // entries per thread of a full-size CTA
    if (T == HPV_THREADS) {
        // every load of the residual in flight before the first use: this sits right behind the wait for the forward
        // kernel, on the step's critical path (four dependent rounds of four loads before)
        float rv[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int idx = tid + q * HPV_THREADS, k = idx >> 6, r = idx & 63;
            rv[q] = (k < nty_e && r < ntx_e) ? a.Res[((size_t)e * a.nty + k) * a.ntx + r] : 0.0f;
        }
#pragma unroll
        for (int q = 0; q < RPT; ++q) s_R[tid + q * HPV_THREADS] = rs * rv[q];
    } else {
        for (int idx = tid; idx < HPV_NP * HPV_NP; idx += T) {
            const int k = idx >> 6, r = idx & 63;
            float v = 0.0f;
            if (k < nty_e && r < ntx_e) v = rs * a.Res[((size_t)e * a.nty + k) * a.ntx + r];
            s_R[idx] = v;
        }
    }
    hpv_sync(c);

    // V_t[r][jl] = sum_k L_t[k][j0+jl] * Rbar[k][r]      (tile: 4 jl x 4 r)
    {
        const int njl4 = (nrows + 3) >> 2;
        const int ntiles = a.n_terms * njl4 * (HPV_NP / 4);
        for (int tile = tid; tile < ntiles; tile += T) {
            const int r4 = tile & 15, rest = tile >> 4;
            const int jl4 = rest % njl4, t = rest / njl4;
            const float* Lt = sm + L.tab[a.terms[t].ltab] + j0 + 4 * jl4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
            for (int k = 0; k < nty_e; ++k) {
                const HpvF4 l4 = hpv_ld4(Lt + k * QP), b4 = hpv_ld4(s_R + k * HPV_NP + 4 * r4);
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ls[i], bs[j], acc[i][j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                HpvF4 o; o.x = acc[0][j]; o.y = acc[1][j]; o.z = acc[2][j]; o.w = acc[3][j];
                hpv_st4(s_V + (t * HPV_NP + 4 * r4 + j) * HPV_ADJ_RSP + 4 * jl4, o);
            }
        }
    }
    hpv_sync(c);

    // Gbar_t[j0+jl][i] = c_t * sum_r V_t[r][jl] * R_t[r][i]      (tile: 4 jl x 4 i)
    {
        const float hwx = a.el_geom[4 * e + 1], hwy = a.el_geom[4 * e + 3];
        const int njl4 = (nrows + 3) >> 2, ni4 = QP >> 2;
        const int ntiles = a.n_terms * njl4 * ni4;
        for (int tile = tid; tile < ntiles; tile += T) {
            const int i4 = tile % ni4, rest = tile / ni4;
            const int jl4 = rest % njl4, t = rest / njl4;
            const float* Vt = s_V + (size_t)t * HPV_NP * HPV_ADJ_RSP + 4 * jl4;
            const float* R = sm + L.tab[a.terms[t].rtab] + 4 * i4;
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
            for (int r = 0; r < ntx_e; ++r) {
                const HpvF4 v4 = hpv_ld4(Vt + r * HPV_ADJ_RSP), w4 = hpv_ld4(R + r * QP);
                const float vs[4] = {v4.x, v4.y, v4.z, v4.w}, ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(vs[i], ws[j], acc[i][j]);
            }
            const float ct = hpv_term_scale(a.terms[t], hwx, hwy);
            float* out = aa.Gbar + ((size_t)t * a.n_el + e) * npts_el;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int jl = 4 * jl4 + i;
                if (jl >= nrows) continue;
                if ((Q & 3) == 0 && 4 * i4 + 3 < Q) {             // rows start at multiples of four floats: one 128-bit store
                    HpvF4 o; o.x = ct * acc[i][0]; o.y = ct * acc[i][1]; o.z = ct * acc[i][2]; o.w = ct * acc[i][3];
                    hpv_st4(out + (j0 + jl) * Q + 4 * i4, o);
                    continue;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int ii = 4 * i4 + j;
                    if (ii < Q) out[(j0 + jl) * Q + ii] = ct * acc[i][j];
                }
            }
        }
    }
    hpv_pdl_trigger();       // at the end, see hpv_varfwd_body (triggering at the start was measured: no gain, r2z2)
}

// ---------------------------------------------------------------------------------------------------------
// K2: reverse sweep through the MLP.
// ---------------------------------------------------------------------------------------------------------
struct HpvBwdArgs {
    HpvVarArgs v;
    const float* Gbar;         // [term][n_points]  (n_points = n_el*rows*Q for the variational loss)
    int n_points;
    // scattered-point mode (boundary / PINN losses): coordinates and per-point adjoints are given directly
    const float* pts;          // [n][dim] or null (then the points are the element quadrature points)
    int stagger_ns;            // start delay per warp "row" (warp / 4), see hpv_mlpbwd_body
};

// Compact layout of a warp's gradient accumulator (the padded parameter layout of hpv_math.cuh without the
// transposed copies):  [W1: DIM x HP][b1: HP] { [W_l: HP x HP][b_l: HP] } x (nhid-1) [Wo: HP][bo, 0, 0, 0]
HPV_HD int hpv_gw_wl(int dim, int hp, int l /*1..nhid-1*/) { return (dim + 1) * hp + (l - 1) * (hp * hp + hp); }
HPV_HD int hpv_gw_wo(int dim, int hp, int nhid) { return (dim + 1) * hp + (nhid - 1) * (hp * hp + hp); }
HPV_HD int hpv_gw_n(int dim, int hp, int nhid) { return hpv_gw_wo(dim, hp, nhid) + hp + 4; }
// padded index -> compact index, or -1 for the transposed copies (which carry no gradient)
HPV_HD int hpv_gw_of_padded(int dim, int hp, int nhid, int ip) {
    const int head = (dim + 1) * hp;
    if (ip < head) return ip;
    const int blk = 2 * hp * hp + hp, j = ip - head, l = j / blk;
    if (l < nhid - 1) {
        const int r = j - l * blk;
        return r < hp * hp + hp ? head + l * (hp * hp + hp) + r : -1;
    }
    return head + (nhid - 1) * (hp * hp + hp) + (j - (nhid - 1) * blk);
}

// Shared-memory plan of the reverse sweep.  Everything but `cst` and `red` is private to a warp (its 32 slot rows
// of every slot, its rows of in0/go, its own gradient accumulator): the warps of a CTA never wait for each other
// inside the sweep.
template <int DIM, int MX, int MY, int HP>
struct HpvBwdSmem {
    typedef HpvMode<DIM, MX, MY> M;
    static constexpr int SP = HpvSP<HP>::value;
    static constexpr int NCH1 = 1 + (M::DX ? 1 : 0) + (M::DY ? 1 : 0);     // channels with a non-zero seed at layer 1
    int cst, slots, go, gw, red, total, NS, slot_sz, gwn;
    HPV_HD HpvBwdSmem(int nhid, int T) {
        int o = 0;
        cst = o; o += 8;                                   // {1,0,0,0} (bias row of the value channel), {0,0,0,0}
        NS = nhid - 1 > 2 ? nhid - 1 : 2;
        slot_sz = M::NCH * T * SP;
        slots = o; o += NS * slot_sz;
        go = o; o += hpv_align4(M::NCH * T);
        gwn = hpv_gw_n(DIM, HP, nhid);
        gw = o; o += (T / 32) * gwn;
        red = o; o += T;
        total = o;
    }
};

// Weight-gradient GEMM over the 32 points of ONE warp:  D[i][j] += sum_ch sum_p IN[ch][p][i] * ADJ[ch][p][j],
// IN and ADJ being the warp's own blocks [channel][32 points][SPI | SPA].
// A lane owns a register tile of 4 rows (i) x TN columns (j), TN = 4 (two packed pairs per row) or 1 (output
// layer: ADJ is one scalar per point and channel).  With fewer tiles than lanes the points (K) are split over
// KSW lanes per tile and the partial tiles are summed with xor-shuffles (fixed order, deterministic); with more
// tiles than lanes (HP = 32) a lane takes several tiles in turn.  The loop body is uniform: a row group beyond
// the rows of IN reads a constant vector instead -- {1,0,0,0} for the bias row in the value channel, zeros
// otherwise -- with stride 0, so there is no per-point branching.  D is this warp's private accumulator in
// shared memory; nothing here synchronises with other warps.
//   KIND 0: hidden layer, D = W[HP][HP], bias row -> b (= W + HP*HP).   KIND 1: output layer, D = Wo[HP] followed
//   by (bo, 0, 0, 0), TN = 1.   KIND 2: first layer, IN rows = (x, y, 1, 0): rows < DIM -> W1, row 2 -> b1.
// NPTS points per call (32: a whole warp tile; 16: half of it -- the tensor-core reverse sweep shares a tile between two
// warps), CHS rows between the channel blocks of IN / ADJ.
template <int SPI, int SPA, int NCH, int NROWS, int NJ, int TN, bool BIAS, int KIND, int HP, int DIM, int NPTS = 32, int CHS = 32>
HPV_HD void hpv_wgrad_warp(const HpvCta& c, const float* IN, const float* ADJ, float* D, float* Db, const float* cst) {
#if defined(HPV_EXP_NO_WGRAD)     // timing experiment only (tools/gpu_r2k.sh): what the kernel costs without the weight gradients
    if (c.tid >= 0) return;
#endif
    constexpr int NI = (NROWS + (BIAS ? 1 : 0) + 3) / 4;
    constexpr int NTILES = NI * NJ;
    constexpr int KSW = NTILES > 16 ? 1 : (NTILES > 8 ? 2 : (NTILES > 4 ? 4 : (NTILES > 2 ? 8 : 16)));
    constexpr int TPB = 32 / KSW;                          // tiles per batch
    constexpr int NP = TN == 4 ? 2 : 1;                    // packed pairs per accumulator row (TN = 1: rows are paired)
    const int lane = c.tid & 31;
    const int ks = lane % KSW;
#pragma unroll 1
    for (int tb = 0; tb < NTILES; tb += TPB) {
        const int tau = tb + lane / KSW;
        const bool active = tau < NTILES;
        const int it = active ? tau / NJ : 0, jt = active ? tau - it * NJ : 0;
        const int r0 = 4 * it;
        const int ty = r0 < NROWS ? 0 : ((BIAS && r0 == NROWS) ? 1 : 2);
        hpv_pair acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = hpv_dup(0.0f); acc[i][1] = hpv_dup(0.0f); }
        if (active) {
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const float* pa = (ty == 0) ? IN + (ch * CHS + ks) * SPI + r0 : ((ty == 1 && ch == 0) ? cst : cst + 4);
                const int sa = (ty == 0) ? KSW * SPI : 0;
                const float* pb = ADJ + (ch * CHS + ks) * SPA + (TN == 4 ? 4 * jt : 0);
                static_assert(NPTS % KSW == 0, "points per call must be a multiple of the K split");
#pragma unroll 8
                for (int k = 0; k < NPTS / KSW; ++k) {
                    const HpvF4 a4 = hpv_ld4(pa);
                    pa += sa;
                    if constexpr (TN == 4) {
                        const HpvF4 b4 = hpv_ld4(pb);
                        const hpv_pair b01 = hpv_pack(b4.x, b4.y), b23 = hpv_pack(b4.z, b4.w);
                        const float as[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const hpv_pair ad = hpv_dup(as[i]);
                            hpv_fma2(acc[i][0], ad, b01);
                            hpv_fma2(acc[i][1], ad, b23);
                        }
                    } else {
                        const hpv_pair bd = hpv_dup(pb[0]);
                        hpv_fma2(acc[0][0], hpv_pack(a4.x, a4.y), bd);
                        hpv_fma2(acc[1][0], hpv_pack(a4.z, a4.w), bd);
                    }
                    pb += KSW * SPA;
                }
            }
        }
        // sum the KSW partial tiles (all lanes take part in the exchanges)
        float v[4][4];
        if constexpr (TN == 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { hpv_unpack(acc[i][0], v[i][0], v[i][1]); hpv_unpack(acc[i][1], v[i][2], v[i][3]); }
        } else {
            hpv_unpack(acc[0][0], v[0][0], v[1][0]);
            hpv_unpack(acc[1][0], v[2][0], v[3][0]);
        }
#pragma unroll
        for (int m = 1; m < KSW; m <<= 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < (TN == 4 ? 4 : 1); ++j) v[i][j] += hpv_shfl_xor(c, v[i][j], m);
        }
        (void)NP;
        if (active && ks == 0) {
            if constexpr (KIND == 1) {
                // rows r0 .. r0+3 of (Wo[0..HP-1], bo, 0, 0, 0) are contiguous
                HpvF4 d = hpv_ld4(D + r0);
                d.x += v[0][0]; d.y += v[1][0]; d.z += v[2][0]; d.w += v[3][0];
                hpv_st4(D + r0, d);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = r0 + i;
                    float* row = nullptr;
                    if (KIND == 0) row = r < NROWS ? D + r * HP : ((BIAS && r == NROWS) ? Db : nullptr);
                    else row = r < DIM ? D + r * HP : (r == 2 ? Db : nullptr);
                    if (row) {
                        HpvF4 d = hpv_ld4(row + 4 * jt);
                        d.x += v[i][0]; d.y += v[i][1]; d.z += v[i][2]; d.w += v[i][3];
                        hpv_st4(row + 4 * jt, d);
                    }
                }
            }
        }
    }
}

#if defined(HPV_EXP_STAMPS) && defined(__CUDACC__)
// Timing experiment only (tools/bwd_stamps.py): %globaltimer of every CTA at entry [0], after the wait for the previous
// kernel [1], at the end of each warp's sweep [2 + warp] and at the CTA's end [20]; hpv_exp_read_bstamps (hpv_k_h20_bwd.cu).
#define HPV_EXP_NBSTAMP 24
static __device__ unsigned long long hpv_exp_bstamps[256 * HPV_EXP_NBSTAMP];
__device__ __forceinline__ void hpv_exp_bstamp(int bid, int slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (bid < 256) hpv_exp_bstamps[bid * HPV_EXP_NBSTAMP + slot] = t;
}
#endif
#if defined(HPV_EXP_STAMPS) && defined(__CUDA_ARCH__)
#define HPV_BSTAMP(cond, slot) do { if (cond) hpv_exp_bstamp(c.bid, (slot)); } while (0)
#else
#define HPV_BSTAMP(cond, slot) do { } while (0)
#endif

// The reverse sweep.  A warp takes a contiguous range of 32-point tiles and runs, per tile and without waiting
// for any other warp: forward recompute (pre-activations kept in the slots), output-layer gradient, then per
// hidden layer top-down the activation adjoint, the weight-gradient GEMM over the warp's points and the
// adjoint product one layer down (transposed weights, the forward product loop), and the first-layer gradient.
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_mlpbwd_body(const HpvCta& c, const HpvBwdArgs& ba) {
    typedef HpvMode<DIM, MX, MY> M;
    typedef HpvState<DIM, MX, MY, HP> State;
    constexpr int SP = HpvSP<HP>::value;
    const HpvVarArgs& a = ba.v;
    const int T = c.nthreads, tid = c.tid, nhid = a.nhid, top = nhid - 1;
    const int lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    const HpvBwdSmem<DIM, MX, MY, HP> L(nhid, T);
    float* sm = reinterpret_cast<float*>(c.smem);
    const float* s_th = HPV_THETA(a.theta_pad);
    float* s_cst = sm + L.cst;
    float* s_gw = sm + L.gw + (size_t)warp * L.gwn;          // this warp's gradient accumulator (compact layout)
    // Every per-point buffer is laid out [warp][channel][32 lanes][...]: a warp addresses its own block with
    // compile-time strides (the slot helpers are called with T = 32, tid = lane).
    constexpr int NCH1 = HpvBwdSmem<DIM, MX, MY, HP>::NCH1;
    constexpr int WSLOT = M::NCH * 32 * SP;
    float* s_go = sm + L.go + warp * (M::NCH * 32);
    float* s_red = sm + L.red;
    // Slots (one row of SP floats per thread and channel).  P(l), l = 1..top-1: pre-activations of hidden layer l
    // kept from the forward recompute, later overwritten by the post-activations (left factor of the W_{l+1}
    // gradient).  X: the running slot -- post-activations feeding the next layer product, then the adjoint of
    // the pre-activations of the layer being differentiated, then (in place) the adjoint handed one layer down.
    // H0: post-activations of the first hidden layer (recomputed), in a slot that is dead by then.
    // With tanh the kept states are "mixed" (hpv_to_mixed): nothing in the reverse part evaluates tanh again
    // except for the first hidden layer, whose state is rebuilt from the coordinates.
    float* const slots_w = sm + L.slots + warp * WSLOT;
    float* const X = slots_w + (size_t)(top >= 1 ? top - 1 : 0) * L.slot_sz;
    float* const H0 = slots_w + (size_t)(top >= 2 ? 0 : 1) * L.slot_sz;
    // layer-1 inputs of the first-layer gradient ([channel][32][4]): written when H0 is dead, in its place
    float* const s_in0 = H0;
    static_assert(NCH1 * 4 <= M::NCH * SP, "the layer-1 input rows must fit into a slot");
#define HPV_P(l) (slots_w + (size_t)((l) - 1) * L.slot_sz)

    HPV_BSTAMP(tid == 0, 0);
    for (int i = lane; i < L.gwn; i += 32) s_gw[i] = 0.0f;
    if (tid < 8) s_cst[tid] = (tid == 0) ? 1.0f : 0.0f;
    hpv_pdl_wait();                                      // Gbar (or the point adjoints) of the previous kernel
    HPV_BSTAMP(tid == 0, 1);
    const float eps = a.eps[0];
    float coef[HPV_MAX_TERMS][HPV_NFIELDS], coef1[HPV_MAX_TERMS][HPV_NFIELDS];      // registers: every loop over them is unrolled
#pragma unroll
    for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
        for (int f = 0; f < HPV_NFIELDS; ++f) {
            coef[t][f] = (t < a.n_terms) ? fmaf(eps, a.terms[t].a1[f], a.terms[t].a0[f]) : 0.0f;
            coef1[t][f] = (t < a.n_terms) ? a.terms[t].a1[f] : 0.0f;
        }
    hpv_sync(c);

    const int Q = a.Q, npts_el = a.rows * Q;
    float deps = 0.0f;
    // Work partition: groups of nwarps consecutive 32-point tiles, a contiguous range of groups per CTA, tile
    // (group, warp) to warp `warp`.  The trip count is the same for every warp of the CTA -- the loop control
    // stays in the uniform datapath, which the constant-memory weight loads of the products below depend on
    // (with per-warp trip counts the compiler falls back to vector-indexed LDC) -- and only the last group of the
    // launch can hold tiles without points (they run with zero adjoints, like the padding points of a tile).
    const long long n_wt = ((long long)ba.n_points + 31) / 32;
    const long long n_grp = (n_wt + nwarps - 1) / nwarps;
    const int grp_begin = (int)(((long long)c.bid * n_grp) / c.nblocks);
    const int grp_end = (int)(((long long)(c.bid + 1) * n_grp) / c.nblocks);
    const float* Wo = s_th + a.off_wo;
    const int g_wo = hpv_gw_wo(DIM, HP, nhid);
#if defined(__CUDA_ARCH__)
    // The warps of a CTA start in lockstep and run the same phase sequence: the three warps that share a scheduler
    // (w, w+4, w+8) would ask for the same pipe at the same time.  A start offset per warp row spreads the phases.
    // (The delay is an operand of nanosleep, not a branch: thread-dependent control flow ahead of the sweep would
    // cost the uniform datapath, see the work partition above.)
    if (ba.stagger_ns > 0) __nanosleep((unsigned)(warp >> 2) * (unsigned)ba.stagger_ns);
#endif

#pragma unroll 1
    for (int grp = grp_begin; grp < grp_end; ++grp) {
        const long long gpl = ((long long)grp * nwarps + warp) * 32 + lane;
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
        // adjoints of the point fields (u, u_x, u_y, u_xx, u_yy)
        float gf[HPV_NFIELDS];
#pragma unroll
        for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = 0.0f;
#pragma unroll
        for (int t = 0; t < HPV_MAX_TERMS; ++t)                        // absent terms: gbar = 0, coef = 0
#pragma unroll
            for (int k = 0; k < HPV_NFIELDS; ++k) gf[k] = fmaf(gbar[t], coef[t][k], gf[k]);
        // directional mode: gx u_x + gy u_y = derivative of u along v = (gx, gy).  The tangent channel is seeded
        // with v, so it carries the scale and its output adjoint is 1.
        const float vx = gf[1], vy = gf[2];
        if constexpr (M::DIR) { gf[1] = 1.0f; gf[2] = 0.0f; }
#define HPV_LAYER1(st)                                                                       \
        do {                                                                                 \
            if constexpr (M::DIR) hpv_layer1_pre_dir<HP>(s_th, x, y, vx, vy, st);            \
            else hpv_layer1_pre<DIM, MX, MY, HP>(s_th, x, y, st);                            \
        } while (0)

        // ---- forward recompute; the pre-activations of hidden layers 1..top-1 stay in their slots ----
        State pre, g;
        HPV_LAYER1(pre);
        int woff = hpv_off_wl(DIM, HP, 1);                           // constant-memory offsets: own induction variables
#pragma unroll 1
        for (int l = 1; l <= top; ++l, woff += 2 * HP * HP + HP) {
            hpv_to_mixed<DIM, MX, MY, HP, ACT>(pre);                     // mixed state of layer l-1
            if (l >= 2) hpv_store_state<DIM, MX, MY, HP>(HPV_P(l - 1), 32, lane, pre);
            hpv_activate_store<DIM, MX, MY, HP, ACT, true>(pre, X, 32, lane);     // h_{l-1}
            const float* W = s_th + woff;
            hpv_matmul_slot<DIM, MX, MY, HP>(W, W + HP * HP, X, 32, lane, pre);
        }
        hpv_to_mixed<DIM, MX, MY, HP, ACT>(pre);                         // pre = mixed state of the top layer from here on
        hpv_activate_store<DIM, MX, MY, HP, ACT, true>(pre, X, 32, lane);         // h_top
        // d loss / d eps needs the fields themselves; the directional mode is only chosen for forms without an
        // eps-dependent coefficient (hpv_form_directional), where this sum is identically 0
        if constexpr (!M::DIR) {
            float f[HPV_NFIELDS];
            hpv_output_slot<DIM, MX, MY, HP>(Wo, X, 32, lane, f);
#pragma unroll
            for (int t = 0; t < HPV_MAX_TERMS; ++t) {
                float d1 = 0.0f;
#pragma unroll
                for (int k = 0; k < HPV_NFIELDS; ++k) d1 = fmaf(coef1[t][k], f[k], d1);
                deps = fmaf(gbar[t], d1, deps);
            }
        }

        // ---- output layer: Wo/bo gradient (left factor h_top in X), adjoint of h_top ----
        s_go[M::C_V * 32 + lane] = gf[0];
        if constexpr (M::DX) s_go[M::C_DX * 32 + lane] = gf[1];
        if constexpr (M::DY) s_go[M::C_DY * 32 + lane] = gf[2];
        if constexpr (M::EX) s_go[M::C_EX * 32 + lane] = gf[3];
        if constexpr (M::EY) s_go[M::C_EY * 32 + lane] = gf[4];
#pragma unroll
        for (int m = 0; m < HP / 2; ++m) {
            const hpv_pair w = hpv_ld_pair(Wo + 2 * m);
            g.v.p[m] = hpv_mul2(hpv_dup(gf[0]), w);
            if constexpr (M::DX) g.dx.p[m] = hpv_mul2(hpv_dup(gf[1]), w);
            if constexpr (M::DY) g.dy.p[m] = hpv_mul2(hpv_dup(gf[2]), w);
            if constexpr (M::EX) g.ex.p[m] = hpv_mul2(hpv_dup(gf[3]), w);
            if constexpr (M::EY) g.ey.p[m] = hpv_mul2(hpv_dup(gf[4]), w);
        }
        hpv_syncwarp(c);
        hpv_wgrad_warp<SP, 1, M::NCH, HP, 1, 1, true, 1, HP, DIM>(c, X, s_go, s_gw + g_wo, nullptr, s_cst);
        hpv_syncwarp(c);

        // ---- hidden layers, top down ----
        int woff_t = a.off_wo - HP * HP;                             // = hpv_off_wt(DIM, HP, top)
#pragma unroll 1
        for (int l = top; l >= 1; --l, woff_t -= 2 * HP * HP + HP) {
            hpv_activate_bwd_store<DIM, MX, MY, HP, ACT, true>(pre, g, X, 32, lane);   // X := adjoint of the pre-activations of layer l
            float* INl = (l - 1 >= 1) ? HPV_P(l - 1) : H0;
            if (l - 1 >= 1) hpv_load_state<DIM, MX, MY, HP>(INl, 32, lane, pre);   // mixed state of layer l-1
            else { HPV_LAYER1(pre); hpv_to_mixed<DIM, MX, MY, HP, ACT>(pre); }
            hpv_activate_store<DIM, MX, MY, HP, ACT, true>(pre, INl, 32, lane);   // h_{l-1}: left factor of the W_l gradient
            hpv_syncwarp(c);
            float* gW = s_gw + hpv_gw_wl(DIM, HP, l);
            hpv_wgrad_warp<SP, SP, M::NCH, HP, HP / 4, 4, true, 0, HP, DIM>(c, INl, X, gW, gW + HP * HP, s_cst);
            hpv_syncwarp(c);
            // adjoint of h_{l-1} = ADJ_l . W_l^T: the forward product loop on the transposed copy, inputs from X
            hpv_matmul_slot<DIM, MX, MY, HP, false>(s_th + woff_t, nullptr, X, 32, lane, g);
        }

        // ---- first layer ----
        hpv_activate_bwd_store<DIM, MX, MY, HP, ACT, true>(pre, g, X, 32, lane);
        {
            HpvF4 o;
            o.x = x; o.y = y; o.z = 1.0f; o.w = 0.0f; hpv_st4(s_in0 + (0 * 32 + lane) * 4, o);
            if constexpr (M::DIR) {                    // the tangent seed is v . W1: its left factor is v
                o.x = vx; o.y = vy; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (1 * 32 + lane) * 4, o);
            } else {
                if constexpr (M::DX) { o.x = 1.0f; o.y = 0.0f; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (M::C_DX * 32 + lane) * 4, o); }
                if constexpr (M::DY) { o.x = 0.0f; o.y = 1.0f; o.z = 0.0f; o.w = 0.0f; hpv_st4(s_in0 + (M::C_DY * 32 + lane) * 4, o); }
            }
        }
        hpv_syncwarp(c);
        // channels that feed W1: value (x, y, 1), d/dx (1, 0, 0), d/dy (0, 1, 0); second-derivative seeds are 0.
        // The IN rows are ordered v, dx, dy, which is also the order of the stored channels (C_V, C_DX, C_DY).
        {
            hpv_wgrad_warp<4, SP, NCH1, 4, HP / 4, 4, false, 2, HP, DIM>(c, s_in0, X, s_gw, s_gw + DIM * HP, s_cst);
            hpv_syncwarp(c);
        }
#undef HPV_LAYER1
    }
#undef HPV_P

    HPV_BSTAMP(lane == 0 && warp < 18, 2 + warp);
    hpv_pdl_trigger();       // the sweep of this CTA is done (at the end, see hpv_varfwd_body; at the start: no gain, r2z3)
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
    HPV_BSTAMP(tid == 0, 20);
}

// K3: grad_pad[i] (+)= sum over CTAs of grad_part[c][i], fixed order.  One CTA = 32 entries x (nthreads / 32)
// groups of partials; a thread sums every ngrp-th partial with four independent accumulators (loads in flight),
// the groups are combined in a fixed order.  accumulate != 0 adds onto the existing value (several losses into
// one gradient).  Returns the reduced value to the threads of group 0 (others: 0).
struct HpvGradReduceArgs {
    const float* grad_part;
    int n_parts, stride, n;    // n entries (theta_pad_n + 1)
    float* grad_pad;
    int accumulate;
};

HPV_HD float hpv_gradreduce_body(const HpvCta& c, const HpvGradReduceArgs& a) {
    hpv_pdl_wait();                                  // the partial gradients of the reverse sweep
    float* s = reinterpret_cast<float*>(c.smem);     // [ngrp][32]
    const int li = c.tid & 31, grp = c.tid >> 5, ngrp = c.nthreads >> 5;
    const int i = c.bid * 32 + li;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    if (i < a.n) {
        const float* src = a.grad_part + i;
        int p = grp;
        for (; p + 3 * ngrp < a.n_parts; p += 4 * ngrp) {
            a0 += src[(size_t)p * a.stride];
            a1 += src[(size_t)(p + ngrp) * a.stride];
            a2 += src[(size_t)(p + 2 * ngrp) * a.stride];
            a3 += src[(size_t)(p + 3 * ngrp) * a.stride];
        }
        for (; p < a.n_parts; p += ngrp) a0 += src[(size_t)p * a.stride];
    }
    s[grp * 32 + li] = (a0 + a1) + (a2 + a3);
    hpv_sync(c);
    float t = 0.0f;
    if (grp == 0 && i < a.n) {
        for (int g2 = 0; g2 < ngrp; ++g2) t += s[g2 * 32 + li];
        if (a.accumulate) t += a.grad_pad[i];
        a.grad_pad[i] = t;
    }
    return t;
}
