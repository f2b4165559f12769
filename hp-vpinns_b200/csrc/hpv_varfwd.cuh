// Fused forward kernel of the variational loss: for every element of the hp-mesh evaluate net_u and the
// needed input derivatives at the tensor Gauss-Lobatto quadrature points, project onto the Jacobi test
// functions by sum factorisation, subtract F_ext and reduce to the element loss and lossv.
// Restates the graph of P2D:68-120, P1D:64-96 and ADI:108-182 (one launch instead of N_el*Nty*Ntx reduce_sum
// sub-graphs).  See DESIGN.md "Forward kernel" for the data flow and the shared-memory plan.
#pragma once
#include "hpv_cta.cuh"
#include "hpv_const.cuh"
#include "hpv_slot.cuh"

// Shared-memory plan.  The region R is time-shared inside a chunk: first the per-thread activation slot of the
// MLP phase, then (re-staged from L2) the test-function tables and the first-contraction output P.
struct HpvFwdSmem {
    int xi1, G, red, flag, R, tab[HPV_NTAB], P, total;   // offsets in floats
    int GS, RMAX;
};

HPV_HD int hpv_tab_mask(const HpvVarArgs& a) {
    int m = 0;
    for (int t = 0; t < a.n_terms; ++t) m |= (1 << a.terms[t].ltab) | (1 << a.terms[t].rtab);
    return m;
}

HPV_HD HpvFwdSmem hpv_fwd_smem(const HpvVarArgs& a, int slot_floats) {
    HpvFwdSmem s;
    int o = 0;
    s.xi1 = o; o += hpv_align4(a.Q);
    s.GS = hpv_align4(HPV_CT * HPV_THREADS + 2 * a.Q);
    int rmax = (HPV_CT * HPV_THREADS) / a.Q + 2;
    s.RMAX = rmax < a.rows ? rmax : a.rows;
    s.G = o; o += a.n_terms * s.GS;
    s.red = o; o += 2 * HPV_THREADS;                    // doubles for the final reduction
    s.flag = o; o += 4;
    s.R = o;
    int p = o;
    int m = hpv_tab_mask(a);
    for (int t = 0; t < HPV_NTAB; ++t) {
        s.tab[t] = -1;
        if (m & (1 << t)) { s.tab[t] = p; p += a.Q * HPV_NP; }
    }
    s.P = p; p += a.n_terms * s.RMAX * HPV_NP;
    const int proj = p - o;
    s.total = o + (proj > slot_floats ? proj : slot_floats);
    return s;
}

HPV_HD float hpv_term_scale(const HpvTerm& t, float hwx, float hwy) {
    float c = t.s;
    if (t.px == 1) c *= hwx; else if (t.px == -1) c /= hwx;
    if (t.py == 1) c *= hwy; else if (t.py == -1) c /= hwy;
    return c;
}

HPV_HD void hpv_stage_tables(const HpvCta& c, const HpvVarArgs& a, float* sm, const int* o_tab) {
    for (int t = 0; t < HPV_NTAB; ++t) {
        if (o_tab[t] < 0) continue;
        const float* src = a.tab[t];
        float* dst = sm + o_tab[t];
        // four loads in flight per thread before the first store (the copy is latency-bound otherwise)
        const int n = a.Q * HPV_NP, step = c.nthreads * 4;
        int i = c.tid * 4;
        for (; i + 3 * step < n; i += 4 * step) {
            const HpvF4 v0 = hpv_ld4(src + i), v1 = hpv_ld4(src + i + step), v2 = hpv_ld4(src + i + 2 * step), v3 = hpv_ld4(src + i + 3 * step);
            hpv_st4(dst + i, v0); hpv_st4(dst + i + step, v1); hpv_st4(dst + i + 2 * step, v2); hpv_st4(dst + i + 3 * step, v3);
        }
        for (; i < n; i += step) hpv_st4(dst + i, hpv_ld4(src + i));
    }
}

template <int DIM, int MX, int MY, int HP, int ACT>
HPV_HD void hpv_varfwd_body(const HpvCta& c, const HpvVarArgs& a) {
    const int T = c.nthreads, tid = c.tid;
    const HpvFwdSmem L = hpv_fwd_smem(a, HpvMode<DIM, MX, MY>::NCH * T * HpvSP<HP>::value);
    float* sm = reinterpret_cast<float*>(c.smem);
    const float* th = HPV_THETA(a.theta_pad);
    float* s_xi1 = sm + L.xi1;
    float* s_G = sm + L.G;
    float* s_P = sm + L.P;
    float* s_slot = sm + L.R;
    float* s_red = sm + L.red;
    int* s_flag = reinterpret_cast<int*>(sm + L.flag);
    const int Q = a.Q;

    for (int i = tid; i < Q; i += T) s_xi1[i] = a.xi1[i];
    const float eps = a.eps[0];
    float coef[HPV_MAX_TERMS][HPV_NFIELDS];              // registers: every loop over them is unrolled
#pragma unroll
    for (int t = 0; t < HPV_MAX_TERMS; ++t)
#pragma unroll
        for (int f = 0; f < HPV_NFIELDS; ++f)
            coef[t][f] = (t < a.n_terms) ? fmaf(eps, a.terms[t].a1[f], a.terms[t].a0[f]) : 0.0f;
    hpv_sync(c);

    const int kt = tid >> 4, rt = tid & 15;              // this thread's 4x4 tile of U (k = 4kt.., r = 4rt..)
    float U[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) U[i][j] = 0.0f;

    const int tpe = a.tiles_per_el;
    const int npts_el = a.rows * Q;
    int t_cur = a.cta_tile_begin[c.bid];
    const int t_end = a.cta_tile_begin[c.bid + 1];

    // The partition counts tiles of a.tile_pts points; a chunk's last pass may leave warps without points.
    const int tile_pts = a.tile_pts, CHUNK_TILES = HPV_CT * HPV_THREADS / tile_pts;
    while (t_cur < t_end) {
        const int e = t_cur / tpe, k0 = t_cur - e * tpe;
        int nt = tpe - k0;
        if (nt > CHUNK_TILES) nt = CHUNK_TILES;
        if (nt > t_end - t_cur) nt = t_end - t_cur;
        const int p0 = k0 * tile_pts;
        int p1 = (k0 + nt) * tile_pts;
        if (p1 > npts_el) p1 = npts_el;
        const int npass = (p1 - p0 + T - 1) / T;
        const int ja = p0 / Q, jb = (p1 - 1) / Q, nrows = jb - ja + 1, base = ja * Q;
        const float lox = a.el_geom[4 * e + 0], hwx = a.el_geom[4 * e + 1];
        const float loy = a.el_geom[4 * e + 2], hwy = a.el_geom[4 * e + 3];

        // (1) clear the field rows of this chunk (first/last row may be only partly covered)
        for (int t = 0; t < a.n_terms; ++t)
            for (int i = tid; i < nrows * Q; i += T) s_G[t * L.GS + i] = 0.0f;
        hpv_sync(c);

        // (2) network and input derivatives at the quadrature points -> projected fields
#pragma unroll 1
        for (int it = 0; it < npass; ++it) {
            const int p = p0 + it * T + tid;
            if (p < p1) {
                const int j = p / Q, i = p - j * Q;
                const float x = fmaf(hwx, s_xi1[i], lox);
                const float y = (DIM == 2) ? fmaf(hwy, s_xi1[j], loy) : 0.0f;
                if (a.field_in) {
                    // projection of a given field (F_ext assembly, P2D:384-414): no network evaluation
                    for (int t = 0; t < a.n_terms; ++t)
                        s_G[t * L.GS + (p - base)] = a.field_in[((size_t)t * a.n_el + e) * npts_el + p];
                } else {
                    float f[HPV_NFIELDS];
                    // the slot rows are private to a thread; laid out per warp so that the strides are immediates
                    hpv_net_point_slot<DIM, MX, MY, HP, ACT>(th, a.nhid, a.off_wo, x, y,
                                                             s_slot + (tid >> 5) * (HpvMode<DIM, MX, MY>::NCH * 32 * HpvSP<HP>::value),
                                                             32, tid & 31, f);
#pragma unroll
                    for (int t = 0; t < HPV_MAX_TERMS; ++t) {
                        float g = 0.0f;
#pragma unroll
                        for (int k = 0; k < HPV_NFIELDS; ++k) g = fmaf(coef[t][k], f[k], g);
                        if (t < a.n_terms) s_G[t * L.GS + (p - base)] = g;
                    }
                }
            }
        }
        hpv_sync(c);

        // the activation slot is dead now: bring the test-function tables into the same region
        hpv_stage_tables(c, a, sm, L.tab);
        hpv_sync(c);

        // (3) first contraction, over the x index:  P_t[jl][r] = c_t * sum_i G_t[jl][i] * R_t[i][r]
        //     item = (term, group of 4 rows, group of 4 test functions): a 4x4 register tile, per i one 128-bit load of
        //     the table row and four scalar loads of the field (16 FMAs per 5 loads; at most one round of items)
        {
            const int ng = (nrows + 3) >> 2;
            const int nitems = a.n_terms * ng * (HPV_NP / 4);
            for (int item = tid; item < nitems; item += T) {
                const int r4 = item & 15, rest = item >> 4;
                const int g4 = rest % ng, t = rest / ng;
                const int jl0 = 4 * g4;
                // rows beyond the chunk read row nrows-1 again (never stored)
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

        t_cur += nt;
        if (t_cur == t_end || t_cur % tpe == 0) {
            // (5) this CTA is done with element e: publish its partial U; the last part to arrive reduces the
            // parts in a fixed order (deterministic), forms the residual and the element loss.
            const int nparts = a.el_nparts[e];
            const int slot = a.el_part_off[e] + (c.bid - a.el_first_cta[e]);
            float* up = a.Upart + (size_t)slot * HPV_NP * HPV_NP;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                HpvF4 o; o.x = U[i][0]; o.y = U[i][1]; o.z = U[i][2]; o.w = U[i][3];
                hpv_st4(up + (4 * kt + i) * HPV_NP + 4 * rt, o);
                U[i][0] = U[i][1] = U[i][2] = U[i][3] = 0.0f;
            }
            hpv_fence();
            hpv_sync(c);
            if (tid == 0) {
                unsigned int prev = hpv_atomic_inc(a.el_done + e);
                s_flag[0] = (prev == (unsigned int)(nparts - 1)) ? 1 : 0;
            }
            hpv_sync(c);
            if (s_flag[0]) {
                hpv_fence();
                const int ntx_e = a.el_ntest[2 * e + 0], nty_e = a.el_ntest[2 * e + 1];
                float S[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) S[i][j] = 0.0f;
                const float* base_p = a.Upart + (size_t)a.el_part_off[e] * HPV_NP * HPV_NP;
#pragma unroll 4
                for (int s = 0; s < nparts; ++s) {               // (unrolled: the L2 loads of several parts in flight)
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        HpvF4 v = hpv_ld4_cg(base_p + (size_t)s * HPV_NP * HPV_NP + (4 * kt + i) * HPV_NP + 4 * rt);
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
                            const float res = S[i][j] - (a.F ? a.F[idx] : 0.0f);
                            a.Res[idx] = res;
                            sq = fmaf(res, res, sq);
                        } else if (k < a.nty && r < a.ntx) {
                            a.Res[((size_t)e * a.nty + k) * a.ntx + r] = 0.0f;
                        }
                    }
                }
                const float tot = hpv_block_sum(c, s_red, sq);
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
                        // (6) every element is finished: lossv = sum of the element losses (P2D:120), fixed order
                        hpv_fence();
                        double* dred = reinterpret_cast<double*>(s_red);
                        double acc = 0.0;
                        for (int i = tid; i < a.n_el; i += T) acc += (double)hpv_ld_cg(a.el_loss + i);
                        dred[tid] = acc;
                        hpv_sync(c);
                        for (int s = T >> 1; s > 0; s >>= 1) {
                            if (tid < s) dred[tid] += dred[tid + s];
                            hpv_sync(c);
                        }
                        if (tid == 0) { a.loss[0] = dred[0]; a.n_done[0] = 0u; }
                    }
                    hpv_sync(c);
                }
            }
        }
        hpv_sync(c);
    }
    // This CTA is done: the next kernel of the step (adjoint projection) may start taking the SM's resources and
    // stage its tables while the remaining CTAs of this grid finish.  The trigger sits at the END of the work on
    // purpose: triggered at the start, the CTAs of a dependent kernel become resident as soon as they fit and then
    // idle at their wait, displacing CTAs of this grid that have not been scheduled yet (measured at C4, r02q:
    // +20 % step time).  This kernel itself is launched with full stream serialisation: its parameters in constant
    // memory were just rewritten by the optimizer, and the constant caches are invalidated at a normal launch.
    hpv_pdl_trigger();
}
