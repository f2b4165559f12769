// Forward kernel of the variational loss, tensor-core form (sm_100a): the same path as hpv_varfwd.cuh --
// net_u and input derivatives at the tensor Gauss-Lobatto points (P2D:81-83,158-185), projection on the test
// functions by sum factorisation (P2D:94-115), Res = U - F_ext, element loss, lossv (P2D:118-120) -- with
//   * the hidden-layer products of the MLP on the 5th-generation tensor cores: tcgen05.mma kind::tf32, M = 128
//     quadrature points, N = 16 / 24 / 32 (output units, padded), K = 8 per instruction; the activations of a layer are the A
//     operand and live in tensor memory (TMEM), written there by the threads that computed them (tcgen05.st); the
//     weights with the bias as an extra input row are K-major B tiles in shared memory; the pre-activations of the
//     next layer accumulate in TMEM and come back with tcgen05.ld.  fp32-class accuracy from TF32 products by the
//     3-term split  A.B ~ Ahi.Bhi + Alo.Bhi + Ahi.Blo  (hi = the TF32 part of the fp32 value, lo = the rest);
//   * the test-function tables staged ONCE per CTA by the bulk-copy engine (TMA, cp.async.bulk + mbarrier) into a
//     region of their own -- the activations no longer need shared memory, so nothing aliases the tables.
// Thread layout of the MLP phase: CTA = 256 threads = 8 warps; warp w works on TMEM sub-partition w % 4 (32 of the
// tile's 128 points, one per lane) and on the units [HP/2 (w / 4), HP/2 (w / 4) + HP/2) of every layer; layer 1 and
// the output layer (K = 2 and N = 1) stay on the FMA pipe.  One elected thread issues the MMAs of a layer, channel
// by channel, each channel committing to its own mbarrier so that the value channel's tanh overlaps the tangent
// channels' products; two CTAs per SM (256 TMEM columns each) overlap one CTA's products with the other's
// activations.  The projection phases are those of hpv_varfwd.cuh.
#pragma once
#include "hpv_cta.cuh"
#include "hpv_slot.cuh"
#include "hpv_umma.cuh"
#include "hpv_varfwd.cuh"

#define HPV_TC_NPAD 32                         // N of the MMA instruction (output units padded)
#define HPV_TC_MTILE 128                       // M: quadrature points per MMA tile

template <int HP> struct HpvTcDims {
    static constexpr int KP = ((HP + 1 + 7) / 8) * 8;          // inputs + bias row, padded to the instruction's K = 8
    static constexpr int HPH = HP / 2;                         // units per thread
    // N of the MMA instruction: the output units rounded up to 8, at least 16.  (N = 24 with M = 128 is outside the
    // N % 16 == 0 rule the PTX manual states for M = 128, but the instruction descriptor carries N >> 3 and the
    // hardware runs it: tools/probes/umma_probe.cu test 3 is exact against the reference at N = 24, at 16.4 instead of
    // 19.9 cycles per instruction; every parity test runs through it.  The tiles in shared and tensor memory keep
    // their 32-row / 32-column pitch.)
    static constexpr int NMMA = HP <= 16 ? 16 : (HP <= 24 ? 24 : HPV_TC_NPAD);
};
// TMEM columns of a CTA: NCH accumulators of NPAD columns, then the A operands (hi, then lo; KP columns per channel).
HPV_HD constexpr int hpv_tc_tmem_need(int nch, int hp) { return nch * HPV_TC_NPAD + 2 * nch * (((hp + 1 + 7) / 8) * 8); }
HPV_HD constexpr int hpv_tc_tmem_cols(int nch, int hp) { return hpv_tc_tmem_need(nch, hp) <= 256 ? 256 : 512; }
HPV_HD constexpr bool hpv_tc_supported(int nch, int hp) { return hpv_tc_tmem_need(nch, hp) <= 512; }

struct HpvFwdTcSmem {
    int xi1, G, red, flag, tab[HPV_NTAB], P, th, part, B, bar, total;   // offsets in floats
    int GS, RMAX, th_n, B_layer;
};

HPV_HD HpvFwdTcSmem hpv_fwd_tc_smem(const HpvVarArgs& a, int dim, int hp, int nch) {
    HpvFwdTcSmem s;
    const int kp = ((hp + 1 + 7) / 8) * 8;
    int o = 0;
    s.xi1 = o; o += hpv_align4(a.Q);
    s.GS = hpv_align4(HPV_CT * HPV_THREADS + 2 * a.Q);
    int rmax = (HPV_CT * HPV_THREADS) / a.Q + 2;
    s.RMAX = rmax < a.rows ? rmax : a.rows;
    s.G = o; o += a.n_terms * s.GS;
    s.red = o; o += 2 * HPV_THREADS;
    s.flag = o; o += 4;
    const int m = hpv_tab_mask(a);
    for (int t = 0; t < HPV_NTAB; ++t) {
        s.tab[t] = -1;
        if (m & (1 << t)) { s.tab[t] = o; o += a.Q * HPV_NP; }
    }
    s.P = o; o += a.n_terms * s.RMAX * HPV_NP;
    s.th_n = hpv_align4((dim + 1) * hp + hp + 4);             // W1, b1, Wo, bo
    s.th = o; o += s.th_n;
    s.part = o; o += nch * HPV_TC_MTILE;
    o = (o + 31) & ~31;                                       // B tiles: 128-byte aligned
    s.B_layer = 2 * kp * HPV_TC_NPAD;                         // hi tile, lo tile
    s.B = o; o += (a.nhid - 1 > 0 ? a.nhid - 1 : 0) * s.B_layer;
    s.bar = o; o += 2 * (HPV_NFIELDS + 1) + 2;                // mbarriers (8 bytes each): one per channel + tables; TMEM base
    s.total = o;
    return s;
}

#if defined(__CUDACC__)

#if defined(HPV_EXP_STAMPS)
// Timing experiment only (tools/gpu_r2q.sh): thread 0 of every CTA of the forward kernel records %globaltimer at the
// phase boundaries and the time it spent in each phase; read back with hpv_exp_read_stamps (hpv_k_h20_fwdtc.cu).
#define HPV_EXP_NSTAMP 12
static __device__ unsigned long long hpv_exp_stamps[1024 * HPV_EXP_NSTAMP];
__device__ __forceinline__ unsigned long long hpv_exp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define HPV_STAMP(i) do { if (tid == 0 && c.bid < 1024) hpv_exp_stamps[c.bid * HPV_EXP_NSTAMP + (i)] = hpv_exp_now(); } while (0)
#define HPV_STAMP_BEGIN(v) unsigned long long v = (tid == 0) ? hpv_exp_now() : 0ull
#define HPV_STAMP_ADD(i, v) do { if (tid == 0 && c.bid < 1024) hpv_exp_stamps[c.bid * HPV_EXP_NSTAMP + (i)] += hpv_exp_now() - v; } while (0)
#else
#define HPV_STAMP(i) do { } while (0)
#define HPV_STAMP_BEGIN(v) do { } while (0)
#define HPV_STAMP_ADD(i, v) do { } while (0)
#endif

__device__ __forceinline__ bool hpv_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// N consecutive TMEM columns of this thread's lane <-> registers (N = HP/2: 4, 10 or 16).
template <int N>
__device__ __forceinline__ void hpv_tmem_ld_n(uint32_t taddr, float* v) {
    static_assert(N % 2 == 0 && N <= 16, "unsupported column count");
    int c = 0;
#pragma unroll
    for (; c + 8 <= N; c += 8) {
        uint32_t r[8];
        hpv_tmem_ld8(taddr + c, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[c + j] = __uint_as_float(r[j]);
    }
    if (N - c >= 4) {
        uint32_t r[4];
        hpv_tmem_ld4(taddr + c, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[c + j] = __uint_as_float(r[j]);
        c += 4;
    }
    if (N - c >= 2) {
        uint32_t r[2];
        hpv_tmem_ld2(taddr + c, r);
        v[c] = __uint_as_float(r[0]); v[c + 1] = __uint_as_float(r[1]);
    }
}
template <int N>
__device__ __forceinline__ void hpv_tmem_st_n(uint32_t taddr, const uint32_t* v) {
    int c = 0;
#pragma unroll
    for (; c + 8 <= N; c += 8) {
        uint32_t r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = v[c + j];
        hpv_tmem_st8(taddr + c, r);
    }
    if (N - c >= 4) { hpv_tmem_st4(taddr + c, v[c], v[c + 1], v[c + 2], v[c + 3]); c += 4; }
    if (N - c >= 2) hpv_tmem_st2(taddr + c, v[c], v[c + 1]);
}

// hi = the TF32 part of x (low 13 mantissa bits cleared), lo = x - hi (exact).  Truncation instead of rounding
// keeps the split at one LOP3 + one FADD; lo then carries up to 13 significant bits of which the tensor core uses
// 11: 2^-21 |x|, the same order as the dropped lo.lo term.
__device__ __forceinline__ void hpv_split_trunc(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// One channel's units of this thread (packed pairs) -> A operand: hi = TF32 part, lo = x - hi, stored with the widest
// tcgen05.st shapes.  (Storing pair by pair with .x2 straight from the pairs, to save the register moves that assemble
// the aligned groups of 8, measured 4 % SLOWER on the forward kernel: 67.0 vs 64.5 us at C3, profiles/round2_tuning.)
template <int HPH>
__device__ __forceinline__ void hpv_tc_split_store(uint32_t addr_hi, uint32_t addr_lo, const hpv_pair* a) {
    uint32_t hi[HPH], lo[HPH];
#pragma unroll
    for (int m = 0; m < HPH / 2; ++m) {
        float h0, h1;
        hpv_unpack(a[m], h0, h1);
        hpv_split_trunc(h0, hi[2 * m], lo[2 * m]);
        hpv_split_trunc(h1, hi[2 * m + 1], lo[2 * m + 1]);
    }
    hpv_tmem_st_n<HPH>(addr_hi, hi);
    hpv_tmem_st_n<HPH>(addr_lo, lo);
}

// The MMAs of one hidden-layer product for all channels, issued by one thread.  TB (TMEM base of this CTA), the
// column offsets and the descriptor strides are compile-time constants and the shared-memory addresses are uniform,
// so every operand sits in a uniform register without a transfer from the vector registers: the instructions go out
// back to back (tools/probes/umma_probe.cu, tests 10 and 13: 17-20 cycles per instruction against 90-190 with
// operands that have to be moved per instruction).
template <uint32_t TB, int NCH, int KP, int NMMA>
__device__ __forceinline__ void hpv_tc_issue_layer(uint32_t bhi, uint32_t blo, uint64_t* bar) {
    constexpr uint32_t colD = 0, colAhi = NCH * HPV_TC_NPAD, colAlo = colAhi + NCH * KP;
    constexpr uint32_t idesc = hpv_umma_idesc_tf32(HPV_TC_MTILE, NMMA, 0, 0);
    constexpr uint32_t LBO = HPV_TC_NPAD * 16, SBO = 128;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int term = 0; term < 3; ++term) {                       // (Alo, Bhi), (Ahi, Blo), (Ahi, Bhi): small terms first
#pragma unroll
            for (int ks = 0; ks < KP / 8; ++ks) {
                const uint64_t bd = hpv_umma_desc((term == 1 ? blo : bhi) + ks * 2 * LBO, LBO, SBO);
                hpv_umma_ts(TB + colD + c * HPV_TC_NPAD, TB + (term == 0 ? colAlo : colAhi) + c * KP + ks * 8, bd, idesc, (term | ks) != 0);
            }
        }
        hpv_umma_commit(&bar[c]);
    }
}


// Scalar form of a layer state for the tensor-core kernels: the values travel between TMEM and the arithmetic as
// plain 32-bit registers (tcgen05.ld/st operate on 32-bit registers; the packed FP32x2 form of hpv_math.cuh costs a
// register move per value to build and take apart the pairs around every TMEM access -- 170 of the 830 instructions
// per thread and tile of the first version, profiles/round2_tuning/varfwd_tc_segments.txt).
template <int DIM, int MX, int MY, int N>
struct HpvStateS {
    typedef HpvMode<DIM, MX, MY> M;
    float v[N];
    float dx[M::DX ? N : 1];
    float dy[M::DY ? N : 1];
    float ex[M::EX ? N : 1];
    float ey[M::EY ? N : 1];
};
template <class M, class S, class F>
__device__ __forceinline__ void hpv_each_ch_s(S& s, F f) {
    f(s.v, M::C_V);
    if constexpr (M::DX) f(s.dx, M::C_DX);
    if constexpr (M::DY) f(s.dy, M::C_DY);
    if constexpr (M::EX) f(s.ex, M::C_EX);
    if constexpr (M::EY) f(s.ey, M::C_EY);
}
//   h = s(z);  dh = s'(z) dz;  d2h = s''(z) dz^2 + s'(z) d2z     (in place)
template <int DIM, int MX, int MY, int N, int ACT>
__device__ __forceinline__ void hpv_activate_s(HpvStateS<DIM, MX, MY, N>& s) {
    typedef HpvMode<DIM, MX, MY> M;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        float a, s1, s2;
        hpv_act<ACT>(s.v[j], a, s1, s2);
        s.v[j] = a;
        if constexpr (M::DX) {
            const float dz = s.dx[j];
            if constexpr (M::EX) s.ex[j] = fmaf(s2 * dz, dz, s1 * s.ex[j]);
            s.dx[j] = s1 * dz;
        }
        if constexpr (M::DY) {
            const float dz = s.dy[j];
            if constexpr (M::EY) s.ey[j] = fmaf(s2 * dz, dz, s1 * s.ey[j]);
            s.dy[j] = s1 * dz;
        }
    }
}
template <int N>
__device__ __forceinline__ void hpv_tc_split_store_s(uint32_t addr_hi, uint32_t addr_lo, const float* h) {
    uint32_t hi[N], lo[N];
#pragma unroll
    for (int j = 0; j < N; ++j) hpv_split_trunc(h[j], hi[j], lo[j]);
    hpv_tmem_st_n<N>(addr_hi, hi);
    hpv_tmem_st_n<N>(addr_lo, lo);
}

// Deterministic block-wide sum with two barriers: xor-shuffle tree inside each warp, then every thread adds the warps'
// sums in warp order (hpv_block_sum's shared-memory tree costs ten barriers; this sits on the critical path of every
// element's last CTA).  `red` holds one float per warp.
__device__ __forceinline__ float hpv_block_sum_warps(const HpvCta& c, float* red, float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    hpv_sync(c);                                   // an earlier use of `red` may still be read
    if ((c.tid & 31) == 0) red[c.tid >> 5] = v;
    hpv_sync(c);
    float r = 0.0f;
    for (int w = 0; w < (c.nthreads >> 5); ++w) r += red[w];
    return r;
}

template <int DIM, int MX, int MY, int HP, int ACT>
__device__ __forceinline__ void hpv_varfwd_tc_body(const HpvCta& c, const HpvVarArgs& a) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int NCH = M::NCH, KP = HpvTcDims<HP>::KP, HPH = HpvTcDims<HP>::HPH;
    constexpr uint32_t TCOLS = (NCH * HPV_TC_NPAD + 2 * NCH * KP) <= 256 ? 256u : 512u;
    static_assert(HPH % 2 == 0, "HP/2 must be even (packed pairs)");
    const int T = c.nthreads, tid = c.tid;
    const HpvFwdTcSmem L = hpv_fwd_tc_smem(a, DIM, HP, NCH);
    float* sm = reinterpret_cast<float*>(c.smem);
    float* s_xi1 = sm + L.xi1;
    float* s_G = sm + L.G;
    float* s_P = sm + L.P;
    float* s_red = sm + L.red;
    float* s_th = sm + L.th;
    float* s_part = sm + L.part;
    uint32_t* s_B = reinterpret_cast<uint32_t*>(sm + L.B);
    int* s_flag = reinterpret_cast<int*>(sm + L.flag);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + L.bar);           // [0..NCH): channels, [HPV_NFIELDS]: tables
    uint32_t* s_tbase = reinterpret_cast<uint32_t*>(s_bar + HPV_NFIELDS + 1);
    const int Q = a.Q, nhid = a.nhid;
    const int warp = tid >> 5, lane = tid & 31, sub = warp & 3, half = warp >> 2, u0 = half * HPH;
    const int prow = sub * 32 + lane;

    // ---- one-time set-up: TMEM, mbarriers, tables by TMA, parameters, weight tiles ----
#if defined(HPV_EXP_STAMPS)
    if (tid == 0 && c.bid < 1024) for (int i = 0; i < HPV_EXP_NSTAMP; ++i) hpv_exp_stamps[c.bid * HPV_EXP_NSTAMP + i] = 0ull;
#endif
    HPV_STAMP(0);
    if (warp == 0) hpv_tmem_alloc(s_tbase, TCOLS);
    if (tid == 0) {
        for (int i = 0; i <= HPV_NFIELDS; ++i) hpv_mbar_init(&s_bar[i], 1);
        hpv_mbar_init_fence();
    }
    __syncthreads();
    if (tid == 0) {
        // the test-function tables: one bulk copy each (the copy engine works while the CTA builds the weight tiles)
        uint32_t bytes = 0;
        for (int t = 0; t < HPV_NTAB; ++t) if (L.tab[t] >= 0) bytes += (uint32_t)(Q * HPV_NP * 4);
        hpv_mbar_expect_tx(&s_bar[HPV_NFIELDS], bytes);
        for (int t = 0; t < HPV_NTAB; ++t)
            if (L.tab[t] >= 0) hpv_bulk_g2s(sm + L.tab[t], a.tab[t], (uint32_t)(Q * HPV_NP * 4), &s_bar[HPV_NFIELDS]);
    }
    for (int i = tid; i < Q; i += T) s_xi1[i] = a.xi1[i];
    {
        // W1 [DIM][HP], b1 [HP] | Wo [HP], bo
        const float* tg = a.theta_pad;
        const int n1 = (DIM + 1) * HP;
        for (int i = tid; i < n1; i += T) s_th[i] = tg[i];
        for (int i = tid; i < HP + 4; i += T) s_th[n1 + i] = tg[a.off_wo + i];
        // B tiles of the hidden layers l = 1 .. nhid-1:  B[n][k] = W_l[k][n] (k < HP), b_l[n] (k = HP), 0 otherwise
        const int per_layer = KP * HPV_TC_NPAD;
        for (int i = tid; i < (nhid - 1) * per_layer; i += T) {
            const int l = i / per_layer, r = i - l * per_layer, n = r / KP, k = r - n * KP;
            const float* W = tg + hpv_off_wl(DIM, HP, l + 1);
            float v = 0.0f;
            if (n < HP) v = k < HP ? W[k * HP + n] : (k == HP ? W[HP * HP + n] : 0.0f);
            uint32_t hi, lo;
            hpv_split_trunc(v, hi, lo);
            const int w = (k >> 2) * (HPV_TC_NPAD * 4) + n * 4 + (k & 3);
            s_B[l * L.B_layer + w] = hi;
            s_B[l * L.B_layer + per_layer + w] = lo;
        }
    }
    const float eps = a.eps[0];
    float coef[HPV_MAX_TERMS][HPV_NFIELDS];
#pragma unroll
    for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
        for (int f = 0; f < HPV_NFIELDS; ++f)
            coef[t][f] = (t < a.n_terms) ? fmaf(eps, a.terms[t].a1[f], a.terms[t].a0[f]) : 0.0f;
    hpv_fence_proxy_async();                   // the weight tiles were written with ordinary stores; the tensor core reads them
    hpv_tc_fence_before();
    __syncthreads();
    hpv_tc_fence_after();
    const uint32_t tb = *s_tbase;
    if (TCOLS == 256u ? (tb != 0u && tb != 256u) : (tb != 0u)) { asm volatile("trap;"); }     // the issue code is specialised on the base
    const uint32_t lane_base = (uint32_t)(sub * 32) << 16;
    constexpr uint32_t colD = 0, colAhi = NCH * HPV_TC_NPAD, colAlo = colAhi + NCH * KP;
    // constant columns of the A operands: input HP of the value channel is the bias input 1, the rest of the padding 0
    if (half == 0) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
            for (int k = HP; k < KP; k += 2) {
                hpv_tmem_st2(tb + lane_base + colAhi + ch * KP + k, (ch == 0 && k == HP) ? __float_as_uint(1.0f) : 0u, 0u);
                hpv_tmem_st2(tb + lane_base + colAlo + ch * KP + k, 0u, 0u);
            }
        }
        hpv_tmem_wait_st();
    }
    HPV_STAMP(1);
    // Programmatic dependent launch: from here on the next kernel of the step (adjoint projection) may be scheduled.  The
    // grid of this kernel is persistent and fully resident, so nothing of it can be displaced (the reason the FFMA form
    // triggers at its end); CTAs of the dependent kernel move into the SMs as CTAs of this grid exit -- the slowest CTA
    // ends 6 us after the median one -- and stage their tables before they wait for this grid to complete
    // (C3 step 214.9 -> 212.1 us, profiles/round2_tuning/pdl_trigger_early.txt).
    hpv_pdl_trigger();
    hpv_mbar_wait(&s_bar[HPV_NFIELDS], 0);      // tables have landed
    HPV_STAMP(2);

    const int kt = tid >> 4, rt = tid & 15;
    float U[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) U[i][j] = 0.0f;

    const int tpe = a.tiles_per_el;
    const int npts_el = a.rows * Q;
    int t_cur = a.cta_tile_begin[c.bid];
    const int t_end = a.cta_tile_begin[c.bid + 1];
    uint32_t phase = 0;
    const float* W1 = s_th;
    const float* b1 = s_th + DIM * HP;
    const float* Wo = s_th + (DIM + 1) * HP;

    const int tile_pts = a.tile_pts, CHUNK_TILES = HPV_CT * HPV_THREADS / tile_pts;
    while (t_cur < t_end) {
        const int e = t_cur / tpe, k0 = t_cur - e * tpe;
        int nt = tpe - k0;
        if (nt > CHUNK_TILES) nt = CHUNK_TILES;
        if (nt > t_end - t_cur) nt = t_end - t_cur;
        const int p0 = k0 * tile_pts;
        int p1 = (k0 + nt) * tile_pts;
        if (p1 > npts_el) p1 = npts_el;
        const int nmt = (p1 - p0 + HPV_TC_MTILE - 1) / HPV_TC_MTILE;            // MMA tiles of 128 points in this chunk
        const int ja = p0 / Q, jb = (p1 - 1) / Q, nrows = jb - ja + 1, base = ja * Q;
        const float lox = a.el_geom[4 * e + 0], hwx = a.el_geom[4 * e + 1];
        const float loy = a.el_geom[4 * e + 2], hwy = a.el_geom[4 * e + 3];

        // (1) clear the field rows of this chunk
        for (int t = 0; t < a.n_terms; ++t)
            for (int i = tid; i < nrows * Q; i += T) s_G[t * L.GS + i] = 0.0f;
        // (no barrier needed here: the first barrier of the MLP phase below orders the clears before the field stores)

        // (2) network and input derivatives at the quadrature points, 128 points at a time
        HPV_STAMP_BEGIN(ts_mlp);
#pragma unroll 1
        for (int mt = 0; mt < nmt; ++mt) {
            const int p = p0 + mt * HPV_TC_MTILE + prow;
            const bool valid = p < p1;
            const int pc = valid ? p : p0;
            const int j = pc / Q, i = pc - j * Q;
            const float x = fmaf(hwx, s_xi1[i], lox);
            const float y = (DIM == 2) ? fmaf(hwy, s_xi1[j], loy) : 0.0f;
            HpvStateS<DIM, MX, MY, HPH> s;
            // first layer, this thread's units: z = b1 + x W1[0,:] (+ y W1[1,:]); dz/dx = W1[0,:]; d2z = 0
#pragma unroll
            for (int jj = 0; jj < HPH; ++jj) {
                const int u = u0 + jj;
                float zz = fmaf(x, W1[u], b1[u]);
                if constexpr (DIM == 2) {
                    zz = fmaf(y, W1[HP + u], zz);
                    if constexpr (M::DY) s.dy[jj] = W1[HP + u];
                }
                s.v[jj] = zz;
                if constexpr (M::DX) s.dx[jj] = W1[u];
                if constexpr (M::EX) s.ex[jj] = 0.0f;
                if constexpr (M::EY) s.ey[jj] = 0.0f;
            }
#pragma unroll 1
            for (int l = 1; l <= nhid; ++l) {
                if (l > 1) {
                    // pre-activations of hidden layer l from TMEM, channel by channel as the products complete
                    hpv_each_ch_s<M>(s, [&](float* zp, int ch) {
#if !defined(HPV_EXP_NO_MMA)          // timing experiment only (tools/gpu_r2l.sh)
                        hpv_mbar_wait(&s_bar[ch], phase);
#endif
                        hpv_tc_fence_after();
                        hpv_tmem_ld_n<HPH>(tb + lane_base + colD + ch * HPV_TC_NPAD + u0, zp);
                        hpv_tmem_wait_ld();
                    });
                    phase ^= 1;
                }
#if !defined(HPV_EXP_NO_ACT)          // timing experiment only (tools/gpu_r2l.sh)
                hpv_activate_s<DIM, MX, MY, HPH, ACT>(s);                // post-activations h, dh, d2h of layer l
#endif
                if (l == nhid) break;
                // split and store as the A operand of the next product
                hpv_each_ch_s<M>(s, [&](float* hp_, int ch) {
                    hpv_tc_split_store_s<HPH>(tb + lane_base + colAhi + ch * KP + u0, tb + lane_base + colAlo + ch * KP + u0, hp_);
                });
                hpv_tmem_wait_st();
                hpv_tc_fence_before();
                __syncthreads();
#if !defined(HPV_EXP_NO_MMA)
                if (warp == 0) {
                    hpv_tc_fence_after();
                    if (hpv_elect_one()) {
                        const uint32_t bhi = hpv_smem_u32(s_B) + (uint32_t)(l - 1) * (uint32_t)(L.B_layer * 4);
                        const uint32_t blo = bhi + (uint32_t)(KP * HPV_TC_NPAD * 4);
                        if (TCOLS == 512u || tb == 0u) hpv_tc_issue_layer<0, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
                        else hpv_tc_issue_layer<256, NCH, KP, HpvTcDims<HP>::NMMA>(bhi, blo, s_bar);
                    }
                    __syncwarp();
                }
#endif
            }
            // output layer: partial sums over this thread's units, combined across the two halves
            {
                float acc[NCH];
                hpv_each_ch_s<M>(s, [&](float* hp_, int ch) {
                    float sacc = 0.0f;
#pragma unroll
                    for (int jj = 0; jj < HPH; ++jj) sacc = fmaf(hp_[jj], Wo[u0 + jj], sacc);
                    acc[ch] = sacc;
                });
                if (half == 1) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) s_part[ch * HPV_TC_MTILE + prow] = acc[ch];
                }
                __syncthreads();
                if (half == 0 && valid) {
#pragma unroll
                    for (int ch = 0; ch < NCH; ++ch) acc[ch] += s_part[ch * HPV_TC_MTILE + prow];
                    float f[HPV_NFIELDS];
                    f[0] = acc[M::C_V] + Wo[HP];
                    f[1] = M::DX ? acc[M::DX ? M::C_DX : 0] : 0.0f;
                    f[2] = M::DY ? acc[M::DY ? M::C_DY : 0] : 0.0f;
                    f[3] = M::EX ? acc[M::EX ? M::C_EX : 0] : 0.0f;
                    f[4] = M::EY ? acc[M::EY ? M::C_EY : 0] : 0.0f;
#pragma unroll
                    for (int t = 0; t < HPV_MAX_TERMS; ++t) {
                        float g = 0.0f;
#pragma unroll
                        for (int k = 0; k < HPV_NFIELDS; ++k) g = fmaf(coef[t][k], f[k], g);
                        if (t < a.n_terms) s_G[t * L.GS + (p - base)] = g;
                    }
                }
                if (nhid == 1) __syncthreads();       // otherwise the next tile's layer barriers protect s_part
            }
        }
        hpv_sync(c);
        HPV_STAMP_ADD(5, ts_mlp);
        HPV_STAMP_BEGIN(ts_p3);

        // (3) first contraction, over the x index:  P_t[jl][r] = c_t * sum_i G_t[jl][i] * R_t[i][r]
#if !defined(HPV_EXP_NO_PROJ)             // timing experiment only (tools/gpu_r2l.sh)
        {
            const int ng = (nrows + 3) >> 2;
            const int nitems = a.n_terms * ng * (HPV_NP / 4);
            for (int item = tid; item < nitems; item += T) {
                const int r4 = item & 15, rest = item >> 4;
                const int g4 = rest % ng, t = rest / ng;
                const int jl0 = 4 * g4;
                const float* g0 = s_G + t * L.GS + (jl0 + 0 < nrows ? jl0 + 0 : nrows - 1) * Q;
                const float* g1 = s_G + t * L.GS + (jl0 + 1 < nrows ? jl0 + 1 : nrows - 1) * Q;
                const float* g2 = s_G + t * L.GS + (jl0 + 2 < nrows ? jl0 + 2 : nrows - 1) * Q;
                const float* g3 = s_G + t * L.GS + (jl0 + 3 < nrows ? jl0 + 3 : nrows - 1) * Q;
                const float* R = sm + L.tab[a.terms[t].rtab] + 4 * r4;
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
                if ((Q & 3) == 0) {
                    // four quadrature nodes per trip: the field rows come in as 128-bit loads (rows start at multiples of
                    // Q floats) -- 8 shared-memory loads per 64 FMAs instead of 20; same summation order as below
#pragma unroll 2
                    for (int i = 0; i < Q; i += 4) {
                        const HpvF4 a0 = hpv_ld4(g0 + i), a1 = hpv_ld4(g1 + i), a2 = hpv_ld4(g2 + i), a3 = hpv_ld4(g3 + i);
                        const float gq[4][4] = {{a0.x, a1.x, a2.x, a3.x}, {a0.y, a1.y, a2.y, a3.y}, {a0.z, a1.z, a2.z, a3.z}, {a0.w, a1.w, a2.w, a3.w}};
#pragma unroll
                        for (int ii = 0; ii < 4; ++ii) {
                            const HpvF4 w = hpv_ld4(R + (i + ii) * HPV_NP);
                            const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                            for (int u = 0; u < 4; ++u)
#pragma unroll
                                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(gq[ii][u], ws[v], acc[u][v]);
                        }
                    }
                } else {
#pragma unroll 4
                    for (int i = 0; i < Q; ++i) {
                        const HpvF4 w = hpv_ld4(R + i * HPV_NP);
                        const float gv[4] = {g0[i], g1[i], g2[i], g3[i]};
                        const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u)
#pragma unroll
                            for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(gv[u], ws[v], acc[u][v]);
                    }
                }
                const float ct = hpv_term_scale(a.terms[t], hwx, hwy);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (jl0 + u < nrows) {
                        HpvF4 o; o.x = ct * acc[u][0]; o.y = ct * acc[u][1]; o.z = ct * acc[u][2]; o.w = ct * acc[u][3];
                        hpv_st4(s_P + (t * L.RMAX + jl0 + u) * HPV_NP + 4 * r4, o);
                    }
                }
            }
        }
        hpv_sync(c);
        HPV_STAMP_ADD(6, ts_p3);
        HPV_STAMP_BEGIN(ts_p4);

        // (4) second contraction, over the y index:  U[k][r] += sum_t sum_jl L_t[ja+jl][k] * P_t[jl][r]
        for (int t = 0; t < a.n_terms; ++t) {
            const float* Lt = sm + L.tab[a.terms[t].ltab] + ja * HPV_NP + 4 * kt;
            const float* Pt = s_P + t * L.RMAX * HPV_NP + 4 * rt;
#pragma unroll 4
            for (int jl = 0; jl < nrows; ++jl) {
                const HpvF4 l4 = hpv_ld4(Lt + jl * HPV_NP), p4 = hpv_ld4(Pt + jl * HPV_NP);
                const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, ps[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) U[i][j] = fmaf(ls[i], ps[j], U[i][j]);
            }
        }
#endif

        HPV_STAMP_ADD(7, ts_p4);
        HPV_STAMP_BEGIN(ts_fin);
        t_cur += nt;
        if (t_cur == t_end || t_cur % tpe == 0) {
            // (5) done with element e: publish the partial U; the last part to arrive reduces the parts in a fixed
            // order, forms the residual and the element loss (as hpv_varfwd_body)
            const int nparts = a.el_nparts[e];
            const int slot = a.el_part_off[e] + (c.bid - a.el_first_cta[e]);
            float* up = a.Upart + (size_t)slot * HPV_NP * HPV_NP;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                HpvF4 o; o.x = U[i][0]; o.y = U[i][1]; o.z = U[i][2]; o.w = U[i][3];
                hpv_st4(up + (4 * kt + i) * HPV_NP + 4 * rt, o);
                U[i][0] = U[i][1] = U[i][2] = U[i][3] = 0.0f;
            }
            // this thread's entries of the right-hand side, fetched while the arrival counter makes its round trip (only
            // the last CTA to arrive uses them: a latency taken off the critical path of the element's last part)
            const int ntx_e = a.el_ntest[2 * e + 0], nty_e = a.el_ntest[2 * e + 1];
            float Fv[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = 4 * kt + i, r = 4 * rt + j;
                    Fv[i][j] = (a.F && k < nty_e && r < ntx_e) ? a.F[((size_t)e * a.nty + k) * a.ntx + r] : 0.0f;
                }
            // release / acquire through thread 0: the barrier orders every thread's stores of the part before thread 0's
            // fence + counter increment, and thread 0's fence after a last arrival before every thread's loads of the
            // other parts (one device-scope fence per CTA instead of one per thread)
            hpv_sync(c);
            if (tid == 0) {
                hpv_fence();
                unsigned int prev = hpv_atomic_inc(a.el_done + e);
                const int last = (prev == (unsigned int)(nparts - 1)) ? 1 : 0;
                if (last) hpv_fence();
                s_flag[0] = last;
            }
            hpv_sync(c);
            if (s_flag[0]) {
                float S[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) S[i][j] = 0.0f;
                const float* base_p = a.Upart + (size_t)a.el_part_off[e] * HPV_NP * HPV_NP;
#pragma unroll 4
                for (int sp = 0; sp < nparts; ++sp) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        HpvF4 v = hpv_ld4_cg(base_p + (size_t)sp * HPV_NP * HPV_NP + (4 * kt + i) * HPV_NP + 4 * rt);
                        S[i][0] += v.x; S[i][1] += v.y; S[i][2] += v.z; S[i][3] += v.w;
                    }
                }
                float sq = 0.0f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = 4 * kt + i;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = 4 * rt + j;
                        if (k < nty_e && r < ntx_e) {
                            const size_t idx = ((size_t)e * a.nty + k) * a.ntx + r;
                            const float res = S[i][j] - Fv[i][j];
                            a.Res[idx] = res;
                            sq = fmaf(res, res, sq);
                        } else if (k < a.nty && r < a.ntx) {
                            a.Res[((size_t)e * a.nty + k) * a.ntx + r] = 0.0f;
                        }
                    }
                }
                const float tot = hpv_block_sum_warps(c, s_red, sq);
                if (tid == 0) {
                    a.el_loss[e] = tot / (float)(ntx_e * nty_e);
                    a.el_done[e] = 0u;
                }
                // the sum over the elements: by the very last CTA, unless the step's loss assembly forms it (defer_total)
                if (!a.defer_total) {
                    if (tid == 0) {
                        hpv_fence();
                        unsigned int prev = hpv_atomic_inc(a.n_done);
                        s_flag[1] = (prev == (unsigned int)(a.n_el - 1)) ? 1 : 0;
                    }
                    hpv_sync(c);
                    if (s_flag[1]) {
                        hpv_fence();
                        double* dred = reinterpret_cast<double*>(s_red);
                        double acc = 0.0;
                        for (int i = tid; i < a.n_el; i += T) acc += (double)hpv_ld_cg(a.el_loss + i);
                        dred[tid] = acc;
                        hpv_sync(c);
                        for (int sft = T >> 1; sft > 0; sft >>= 1) {
                            if (tid < sft) dred[tid] += dred[tid + sft];
                            hpv_sync(c);
                        }
                        if (tid == 0) { a.loss[0] = dred[0]; a.n_done[0] = 0u; }
                    }
                    hpv_sync(c);
                }
            }
        }
        hpv_sync(c);
        HPV_STAMP_ADD(8, ts_fin);
    }
    HPV_STAMP(3);
    hpv_tc_fence_before();
    __syncthreads();
    if (warp == 0) hpv_tmem_dealloc(tb, TCOLS);
    HPV_STAMP(4);
}

#endif  // __CUDACC__
