// Layer products of the MLP with the weights read from constant memory through the uniform datapath (see
// hpv_const.cuh), the activations of a point kept in a per-thread row of shared memory
// ("slot": [channel][thread][unit], unit contiguous, row stride SP chosen so that 128-bit accesses of a warp
// are bank-conflict free) and the accumulation done with packed FP32x2 fused multiply-adds (fma.rn.f32x2,
// SASS FFMA2): one issue slot per two FMAs, which leaves the other issue slot of the FMA pipe's two-cycle
// occupancy to the loads, the activation arithmetic and the loop control.  The loops over the input units are
// rolled (4 units per trip), so the hot code stays a few KB and resident in the instruction cache -- the fully
// unrolled register-resident form of hpv_math.cuh stalls on instruction fetch (profiles/r01a_*).
#pragma once
#include "hpv_math.cuh"

template <int HP> struct HpvSP { static constexpr int value = ((HP / 4) % 2 == 1) ? HP : HP + 4; };
HPV_HD int hpv_sp(int hp) { return ((hp / 4) % 2 == 1) ? hp : hp + 4; }
// channels of a (canonical) derivative mode: value + one per carried first / second derivative
HPV_HD int hpv_mode_nch(int dim, int mx, int my) {
    int n = 1 + (mx >= 1 ? 1 : 0) + (mx >= 2 ? 1 : 0);
    if (dim == 2) n += (my >= 1 ? 1 : 0) + (my >= 2 ? 1 : 0);
    return n;
}
HPV_HD int hpv_slot_floats(int dim, int mx, int my, int hp, int T) { return hpv_mode_nch(dim, mx, my) * T * hpv_sp(hp); }

// Visit the channels a mode carries: f(array of HP/2 packed pairs, channel slot index).
template <class M, class S, class F>
HPV_HD void hpv_each_ch(S& s, F f) {
    f(s.v.p, M::C_V);
    if constexpr (M::DX) f(s.dx.p, M::C_DX);
    if constexpr (M::DY) f(s.dy.p, M::C_DY);
    if constexpr (M::EX) f(s.ex.p, M::C_EX);
    if constexpr (M::EY) f(s.ey.p, M::C_EY);
}

// 128-bit shared-memory access of two packed pairs (4 consecutive units of a slot row).
HPV_HD void hpv_st_pairs(float* p, hpv_pair a, hpv_pair b) {
#if defined(__CUDA_ARCH__)
    ulonglong2 v; v.x = a; v.y = b;
    *reinterpret_cast<ulonglong2*>(p) = v;
#else
    p[0] = a.lo; p[1] = a.hi; p[2] = b.lo; p[3] = b.hi;
#endif
}
HPV_HD void hpv_ld_pairs(const float* p, hpv_pair& a, hpv_pair& b) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    a = v.x; b = v.y;
#else
    a.lo = p[0]; a.hi = p[1]; b.lo = p[2]; b.hi = p[3];
#endif
}

template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_store_state(float* slot, int T, int tid, const HpvState<DIM, MX, MY, HP>& s) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value;
    hpv_each_ch<M>(s, [&](const hpv_pair* a, int c) {
        float* row = slot + ((size_t)c * T + tid) * SP;
        // (64-bit stores straight from the packed pairs were measured no faster: two-way bank conflicts, r02d)
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) hpv_st_pairs(row + 4 * j4, a[2 * j4], a[2 * j4 + 1]);
    });
}

// Store the channels of two pair results (units 4*j4 .. 4*j4+3) with one 128-bit store per channel.
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_store_quad(float* slot, int T, int tid, int j4, const HpvPairOut<DIM, MX, MY>& a, const HpvPairOut<DIM, MX, MY>& b) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value;
    float* row = slot + (size_t)tid * SP + 4 * j4;
    hpv_st_pairs(row + (size_t)M::C_V * T * SP, a.v, b.v);
    if constexpr (M::DX) hpv_st_pairs(row + (size_t)M::C_DX * T * SP, a.dx, b.dx);
    if constexpr (M::DY) hpv_st_pairs(row + (size_t)M::C_DY * T * SP, a.dy, b.dy);
    if constexpr (M::EX) hpv_st_pairs(row + (size_t)M::C_EX * T * SP, a.ex, b.ex);
    if constexpr (M::EY) hpv_st_pairs(row + (size_t)M::C_EY * T * SP, a.ey, b.ey);
}

// hpv_activate followed by hpv_store_state of the result, quad by quad: the post-activations go straight from the
// arithmetic into the store operands and never exist as a whole state in registers.  s is left unchanged.
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED>
HPV_HD void hpv_activate_store(const HpvState<DIM, MX, MY, HP>& s, float* slot, int T, int tid) {
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvPairOut<DIM, MX, MY> a, b;
        hpv_activate_pair<DIM, MX, MY, HP, ACT, MIXED>(s, 2 * j4, a);
        hpv_activate_pair<DIM, MX, MY, HP, ACT, MIXED>(s, 2 * j4 + 1, b);
        hpv_store_quad<DIM, MX, MY, HP>(slot, T, tid, j4, a, b);
    }
}

// hpv_activate_bwd followed by hpv_store_state of the resulting adjoints, quad by quad (g is left unchanged; it is
// dead afterwards in the reverse sweep: the next product overwrites it).
template <int DIM, int MX, int MY, int HP, int ACT, bool MIXED>
HPV_HD void hpv_activate_bwd_store(const HpvState<DIM, MX, MY, HP>& z, const HpvState<DIM, MX, MY, HP>& g, float* slot, int T, int tid) {
#pragma unroll
    for (int j4 = 0; j4 < HP / 4; ++j4) {
        HpvPairOut<DIM, MX, MY> a, b;
        hpv_activate_bwd_pair<DIM, MX, MY, HP, ACT, MIXED>(z, g, 2 * j4, a);
        hpv_activate_bwd_pair<DIM, MX, MY, HP, ACT, MIXED>(z, g, 2 * j4 + 1, b);
        hpv_store_quad<DIM, MX, MY, HP>(slot, T, tid, j4, a, b);
    }
}

template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_load_state(const float* slot, int T, int tid, HpvState<DIM, MX, MY, HP>& s) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value;
    hpv_each_ch<M>(s, [&](hpv_pair* a, int c) {
        const float* row = slot + ((size_t)c * T + tid) * SP;
#pragma unroll
        for (int j4 = 0; j4 < HP / 4; ++j4) hpv_ld_pairs(row + 4 * j4, a[2 * j4], a[2 * j4 + 1]);
    });
}

// out = in . W (+ b on the value channel), inputs read from this thread's slot rows, outputs in registers.
// BIAS = false (b unused) gives the plain product, used with the transposed weights for the adjoint sweep.
template <int DIM, int MX, int MY, int HP, bool BIAS = true>
HPV_HD void hpv_matmul_slot(const float* W, const float* b, const float* slot, int T, int tid,
                            HpvState<DIM, MX, MY, HP>& out) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value, NCH = M::NCH;
    hpv_pair acc[NCH][HP / 2];
#pragma unroll
    for (int m = 0; m < HP / 2; ++m) {
        acc[0][m] = BIAS ? hpv_ld_pair(b + 2 * m) : hpv_dup(0.0f);
#pragma unroll
        for (int c = 1; c < NCH; ++c) acc[c][m] = hpv_dup(0.0f);
    }
    static_assert(HP % 4 == 0, "padded hidden width must be a multiple of 4 (128-bit weight loads)");
    // two independent induction variables: the weight pointer (constant memory, must stay in uniform registers so
    // that the loads are LDCU and the FFMA2 take a UR operand) and the slot-row pointer (per-thread, vector)
    // The inputs of trip i4 + 1 are fetched while trip i4 computes (the loop is rolled, so nothing else would hide
    // the shared-memory latency at the top of each trip); the last trip re-reads its own inputs.
    const float* row = slot + (size_t)tid * SP;
    const float* wr0 = W;
    HpvF4 xn[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) xn[c] = hpv_ld4(row + (size_t)c * T * SP);
#if defined(HPV_EXP_NO_PROD)      // timing experiment only (tools/gpu_r2k.sh): what the kernel costs without the products
    if (tid >= 0) return;
#endif
#pragma unroll 1
    for (int i4 = 0; i4 < HP / 4; ++i4, wr0 += 4 * HP) {
        float x[NCH][4];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { x[c][0] = xn[c].x; x[c][1] = xn[c].y; x[c][2] = xn[c].z; x[c][3] = xn[c].w; }
        if (i4 + 1 < HP / 4) row += 4;
#pragma unroll
        for (int c = 0; c < NCH; ++c) xn[c] = hpv_ld4(row + (size_t)c * T * SP);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float* wr = wr0 + k * HP;
            hpv_pair xd[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) xd[c] = hpv_dup(x[c][k]);
#pragma unroll
            for (int m4 = 0; m4 < HP / 4; ++m4) {
                // four weights per load (constant memory: one LDCU.128 feeding 2 * NCH FFMA2); every offset of the
                // padded layout is a multiple of 4 floats
                hpv_pair w0, w1;
                hpv_ld_pairs(wr + 4 * m4, w0, w1);
#pragma unroll
                for (int c = 0; c < NCH; ++c) { hpv_fma2(acc[c][2 * m4], xd[c], w0); hpv_fma2(acc[c][2 * m4 + 1], xd[c], w1); }
            }
        }
    }
    hpv_each_ch<M>(out, [&](hpv_pair* a, int c) {
#pragma unroll
        for (int m = 0; m < HP / 2; ++m) a[m] = acc[c][m];
    });
}

// Output layer from the slot: fields (u, u_x, u_y, u_xx, u_yy); absent ones are 0.
template <int DIM, int MX, int MY, int HP>
HPV_HD void hpv_output_slot(const float* Wo, const float* slot, int T, int tid, float f[HPV_NFIELDS]) {
    typedef HpvMode<DIM, MX, MY> M;
    constexpr int SP = HpvSP<HP>::value, NCH = M::NCH;
    float acc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) acc[c] = 0.0f;
    const float* row = slot + (size_t)tid * SP;
#pragma unroll 1
    for (int i4 = 0; i4 < HP / 4; ++i4) {
        const float w0 = Wo[4 * i4], w1 = Wo[4 * i4 + 1], w2 = Wo[4 * i4 + 2], w3 = Wo[4 * i4 + 3];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const HpvF4 v = hpv_ld4(row + (size_t)c * T * SP + 4 * i4);
            acc[c] = fmaf(v.x, w0, acc[c]); acc[c] = fmaf(v.y, w1, acc[c]);
            acc[c] = fmaf(v.z, w2, acc[c]); acc[c] = fmaf(v.w, w3, acc[c]);
        }
    }
    f[0] = acc[M::C_V] + Wo[HP];
    f[1] = M::DX ? acc[M::DX ? M::C_DX : 0] : 0.0f;
    f[2] = M::DY ? acc[M::DY ? M::C_DY : 0] : 0.0f;
    f[3] = M::EX ? acc[M::EX ? M::C_EX : 0] : 0.0f;
    f[4] = M::EY ? acc[M::EY ? M::C_EY : 0] : 0.0f;
}

// Whole network at one point through a slot (the forward kernel's inner loop).
template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_net_point_slot(const float* th, int nhid, int off_wo, float x, float y, float* slot, int T, int tid,
                               float f[HPV_NFIELDS]) {
    HpvState<DIM, MX, MY, HP> s;
    hpv_layer1_pre<DIM, MX, MY, HP>(th, x, y, s);
    hpv_activate<DIM, MX, MY, HP, ACT>(s);
    hpv_store_state<DIM, MX, MY, HP>(slot, T, tid, s);
#pragma unroll 1
    for (int l = 1; l < nhid; ++l) {
        const float* W = th + hpv_off_wl(DIM, HP, l);
        hpv_matmul_slot<DIM, MX, MY, HP>(W, W + HP * HP, slot, T, tid, s);
        hpv_activate<DIM, MX, MY, HP, ACT>(s);        // (the fused hpv_activate_store measured 3 % slower here, r02n)
        hpv_store_state<DIM, MX, MY, HP>(slot, T, tid, s);
    }
    hpv_output_slot<DIM, MX, MY, HP>(th + off_wo, slot, T, tid, f);
}
