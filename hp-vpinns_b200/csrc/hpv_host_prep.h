// Host-side preparation shared by the C-ABI (hpv_api.cu) and the thread-emulation harness (tests/emu):
// padded parameter layout, weighted/transposed test-function tables, the projected terms of each variational
// form, and the static work partition of the forward kernel.  Plain C++, no CUDA.
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "hpv_math.cuh"

enum { HPV_POISSON1D = 0, HPV_POISSON2D = 1, HPV_ADVDIFF = 2 };

struct HpvNet {
    int dim = 0, nhid = 0, hp = 0, act = 0, n_theta = 0, theta_pad_n = 0;
    std::vector<int> layers;
    std::vector<int> pad_index;        // reference order (per layer: W row-major [in][out], then b) -> padded index
    std::vector<int> pad_index2;       // second padded location of the same parameter (transposed copy) or -1
};

inline bool hpv_net_setup(HpvNet& n, int dim, const int* layers, int n_layers, int act, std::string& err) {
    if (dim != 1 && dim != 2) { err = "dim must be 1 or 2"; return false; }
    if (n_layers < 3) { err = "need at least one hidden layer"; return false; }
    if (layers[0] != dim) { err = "layers[0] must equal dim"; return false; }
    if (layers[n_layers - 1] != 1) { err = "the network output must be scalar (layers[-1] == 1)"; return false; }
    if (act != HPV_ACT_SIN && act != HPV_ACT_TANH) { err = "activation must be 0 (sin) or 1 (tanh)"; return false; }
    int nhid = n_layers - 2, H = 0;
    if (nhid > HPV_MAX_HIDDEN) { err = "too many hidden layers (max 8)"; return false; }
    for (int l = 1; l <= nhid; ++l) {
        if (layers[l] < 1) { err = "hidden width must be positive"; return false; }
        if (layers[l] > H) H = layers[l];
    }
    int hp = H <= 8 ? 8 : (H <= 20 ? 20 : (H <= 32 ? 32 : 0));
    if (!hp) { err = "hidden width > 32 is not supported by the fused kernels"; return false; }
    n.dim = dim; n.nhid = nhid; n.hp = hp; n.act = act;
    n.layers.assign(layers, layers + n_layers);
    n.theta_pad_n = hpv_theta_pad_n(dim, hp, nhid);
    n.pad_index.clear();
    n.pad_index2.clear();
    for (int l = 0; l < n_layers - 1; ++l) {
        const int in = layers[l], out = layers[l + 1];
        int offW, offb, stride;
        if (l == 0) { offW = hpv_off_w1(); offb = hpv_off_b1(dim, hp); stride = hp; }
        else if (l < nhid) { offW = hpv_off_wl(dim, hp, l); offb = offW + hp * hp; stride = hp; }
        else { offW = hpv_off_wo(dim, hp, nhid); offb = offW + hp; stride = 1; }
        for (int i = 0; i < in; ++i)
            for (int j = 0; j < out; ++j) {
                n.pad_index.push_back(offW + i * stride + j);
                n.pad_index2.push_back((l >= 1 && l < nhid) ? hpv_off_wt(dim, hp, l) + j * hp + i : -1);
            }
        for (int j = 0; j < out; ++j) { n.pad_index.push_back(offb + j); n.pad_index2.push_back(-1); }
    }
    n.n_theta = (int)n.pad_index.size();
    return true;
}

inline void hpv_pad_theta(const HpvNet& n, const double* theta, std::vector<float>& out) {
    out.assign(n.theta_pad_n, 0.0f);
    for (int i = 0; i < n.n_theta; ++i) {
        out[n.pad_index[i]] = (float)theta[i];
        if (n.pad_index2[i] >= 0) out[n.pad_index2[i]] = (float)theta[i];
    }
}

// Instantiated derivative modes: 1-D (mx); 2-D (0,0), (1,1), (2,1), (2,2).
inline void hpv_canon_mode(int dim, int& mx, int& my) {
    if (dim == 1) { my = 0; return; }
    if (mx == 0 && my == 0) return;
    if (mx <= 1 && my <= 1) { mx = 1; my = 1; return; }
    if (mx == 2 && my <= 1) { my = 1; return; }
    mx = 2; my = 2;
}

inline void hpv_mode_of_coef(int dim, const float* a0, const float* a1, int& mx, int& my) {
    auto nz = [&](int f) { return a0[f] != 0.0f || a1[f] != 0.0f; };
    if (nz(3) && mx < 2) mx = 2;
    if (nz(1) && mx < 1) mx = 1;
    if (dim == 2) {
        if (nz(4) && my < 2) my = 2;
        if (nz(2) && my < 1) my = 1;
    }
}

struct HpvForm {
    int n_terms = 0;
    HpvTerm terms[HPV_MAX_TERMS];
    bool has_rhs = true;               // Res = U - F_ext (Poisson) or Res = U (AdvDiff, ADI:180)
    bool fold_boundary = false;        // P1D var_form 3: boundary terms folded into table 2
    int mx = 0, my = 0;
};

inline HpvTerm hpv_term_zero() {
    HpvTerm t;
    for (int f = 0; f < HPV_NFIELDS; ++f) { t.a0[f] = 0.0f; t.a1[f] = 0.0f; }
    t.s = 1.0f; t.px = 0; t.py = 0; t.ltab = 0; t.rtab = 0;
    return t;
}

// Field indices: 0 u, 1 u_x, 2 u_y (u_t), 3 u_xx, 4 u_yy.  Tables: 0 T*w, 1 D1*w, 2 D2*w, 3 ONE.
// J = Jx*Jy with Jx, Jy the element half-widths (P2D:77-79), so J/Jx = Jy and J/Jy = Jx.
inline bool hpv_form_setup(HpvForm& fm, int problem, int var_form, double V, std::string& err) {
    fm = HpvForm();
    HpvTerm a = hpv_term_zero(), b = hpv_term_zero();
    if (problem == HPV_POISSON2D) {
        if (var_form == 0) {            // P2D:94-96   J * sum wx phi_r wy phi_k (u_xx + u_yy)
            a.a0[3] = 1; a.a0[4] = 1; a.px = 1; a.py = 1; a.ltab = 0; a.rtab = 0; fm.n_terms = 1;
        } else if (var_form == 1) {     // P2D:99-105  -(J/Jx) sum wx phi'_r wy phi_k u_x - (J/Jy) sum wx phi_r wy phi'_k u_y
            a.a0[1] = 1; a.s = -1; a.px = 0; a.py = 1; a.ltab = 0; a.rtab = 1;
            b.a0[2] = 1; b.s = -1; b.px = 1; b.py = 0; b.ltab = 1; b.rtab = 0; fm.n_terms = 2;
        } else if (var_form == 2) {     // P2D:109-115 J sum wx phi''_r wy phi_k u + J sum wx phi_r wy phi''_k u
            a.a0[0] = 1; a.px = 1; a.py = 1; a.ltab = 0; a.rtab = 2;
            b.a0[0] = 1; b.px = 1; b.py = 1; b.ltab = 2; b.rtab = 0; fm.n_terms = 2;
        } else { err = "Poisson-2D var_form must be 0, 1 or 2"; return false; }
    } else if (problem == HPV_POISSON1D) {
        a.ltab = 3;
        if (var_form == 1) {            // P1D:83-84   -J sum w phi_i u''
            a.a0[3] = 1; a.s = -1; a.px = 1; a.rtab = 0;
        } else if (var_form == 2) {     // P1D:86-87   sum w phi'_i u'
            a.a0[1] = 1; a.rtab = 1;
        } else if (var_form == 3) {     // P1D:89-91   -(1/J) sum w phi''_i u + (1/J)(u(b) phi'_i(1) - u(a) phi'_i(-1))
            a.a0[0] = 1; a.px = -1; a.rtab = 2; fm.fold_boundary = true;
        } else { err = "Poisson-1D var_form must be 1, 2 or 3"; return false; }
        fm.n_terms = 1;
    } else if (problem == HPV_ADVDIFF) {
        fm.has_rhs = false;
        if (var_form == 0) {            // ADI:162-167 J sum wx phi_r wt phi_k (u_t + V u_x - eps u_xx)
            a.a0[2] = 1; a.a0[1] = (float)V; a.a1[3] = -1; a.px = 1; a.py = 1; a.ltab = 0; a.rtab = 0; fm.n_terms = 1;
        } else if (var_form == 1) {     // ADI:171-174 J sum (u_t + V u_x) phi phi + eps (J/Jx) sum wx phi'_r wt phi_k u_x
            a.a0[2] = 1; a.a0[1] = (float)V; a.px = 1; a.py = 1; a.ltab = 0; a.rtab = 0;
            b.a1[1] = 1; b.px = 0; b.py = 1; b.ltab = 0; b.rtab = 1; fm.n_terms = 2;
        } else { err = "AdvDiff var_form must be 0 or 1"; return false; }
    } else { err = "unknown problem"; return false; }
    fm.terms[0] = a; fm.terms[1] = b;
    const int dim = (problem == HPV_POISSON1D) ? 1 : 2;
    for (int t = 0; t < fm.n_terms; ++t) hpv_mode_of_coef(dim, fm.terms[t].a0, fm.terms[t].a1, fm.mx, fm.my);
    hpv_canon_mode(dim, fm.mx, fm.my);
    return true;
}

// The reverse sweep may carry ONE tangent along the per-point direction (gx, gy) instead of the x and y tangents
// (HpvMode<2,1,0>::DIR) when the form projects first derivatives only and no coefficient depends on eps (the
// eps gradient needs u_x, u_y separately): Poisson-2D var_form 1 (P2D:99-105) -- the headline configuration.
inline bool hpv_form_directional(int dim, const HpvForm& fm) {
    if (dim != 2 || fm.mx != 1 || fm.my != 1) return false;
    for (int t = 0; t < fm.n_terms; ++t)
        for (int f = 0; f < HPV_NFIELDS; ++f)
            if (fm.terms[t].a1[f] != 0.0f) return false;
    return true;
}

// Transposed, quadrature-weighted tables [Q][HPV_NP]:  tab[q][n] = table[n][q] * w[q].
// T, D1, D2 are the reference's Test_fcn / dTest_fcn outputs on the 1-D nodes, row-major [N][Q].
// d1_bound [N][2] = phi'_n(-1), phi'_n(+1) (only for the folded boundary table of P1D var_form 3, which
// relies on the Gauss-LOBATTO nodes containing both end points: xi[0] = -1, xi[Q-1] = +1).
inline void hpv_build_tables(int Q, int N, const double* w, const double* T, const double* D1, const double* D2,
                             const double* d1_bound, bool fold_boundary, std::vector<float> tabs[HPV_NTAB]) {
    const double* src[3] = {T, D1, D2};
    for (int t = 0; t < 3; ++t) {
        tabs[t].assign((size_t)Q * HPV_NP, 0.0f);
        if (!src[t]) continue;
        for (int n = 0; n < N; ++n)
            for (int q = 0; q < Q; ++q) {
                double v = src[t][(size_t)n * Q + q] * w[q];
                if (t == 2 && fold_boundary) {
                    v = -v;
                    if (q == 0) v -= d1_bound[2 * n + 0];
                    if (q == Q - 1) v += d1_bound[2 * n + 1];
                }
                tabs[t][(size_t)q * HPV_NP + n] = (float)v;
            }
    }
    tabs[3].assign((size_t)Q * HPV_NP, 0.0f);
    tabs[3][0] = 1.0f;
}

// Natural-layout copies [HPV_NP][QP] (QP = Q rounded up to a multiple of 4, 4 floats of tail padding) for the
// adjoint projection, whose contractions run along the other index.
inline void hpv_natural_tables(int Q, const std::vector<float> tabs[HPV_NTAB], std::vector<float> nat[HPV_NTAB]) {
    const int QP = (Q + 3) & ~3;
    for (int t = 0; t < HPV_NTAB; ++t) {
        nat[t].assign((size_t)HPV_NP * QP + 4, 0.0f);
        for (int q = 0; q < Q; ++q)
            for (int n = 0; n < HPV_NP; ++n) nat[t][(size_t)n * QP + q] = tabs[t][(size_t)q * HPV_NP + n];
    }
}

struct HpvPartition {
    int tiles_per_el = 0, n_ctas = 0, total_parts = 0;
    std::vector<int> cta_tile_begin, el_first_cta, el_part_off, el_nparts;
};

// Cost model of a CTA's share for the weighted partition (all in units of one tile's run time): a tile costs 1 on the
// first `first_wave` CTAs and `wave2` on the others (the second CTA an SM takes gets fewer issue slots than the first),
// every element boundary inside a CTA's range costs `cross` (one more chunk, one more partial-U publication, and such a
// CTA is usually the last to arrive at the element it finishes).  cross = 0, wave2 = 1: equal tile counts.
struct HpvPartitionCost {
    double cross = 0.0;
    int first_wave = 0;
    double wave2 = 1.0;
};

// Tiles of tile_pts points (never straddling elements), a contiguous range of tiles per CTA; a CTA gets at least
// cta_pts points' worth of tiles (one full pass of its threads) unless there is less work than that.  With a cost
// model the ranges are chosen so that the modelled costs of the CTAs are (nearly) equal.
inline void hpv_partition(HpvPartition& p, int n_el, int pts_per_el, int tile_pts, int max_ctas, int cta_pts = 0,
                          const HpvPartitionCost* cost = nullptr) {
    p.tiles_per_el = (pts_per_el + tile_pts - 1) / tile_pts;
    const long long ntiles = (long long)n_el * p.tiles_per_el;
    const long long per_cta = cta_pts > tile_pts ? cta_pts / tile_pts : 1;
    const long long want = (ntiles + per_cta - 1) / per_cta;
    p.n_ctas = (int)(want < max_ctas ? want : max_ctas);
    if (p.n_ctas < 1) p.n_ctas = 1;
    p.cta_tile_begin.resize(p.n_ctas + 1);
    for (int c = 0; c <= p.n_ctas; ++c) p.cta_tile_begin[c] = (int)((ntiles * c) / p.n_ctas);
    if (cost && (cost->cross > 0.0 || cost->wave2 != 1.0) && ntiles > p.n_ctas) {
        const int n = p.n_ctas, tpe = p.tiles_per_el;
        // Greedy in CTA order with a target that is re-derived from what is left (so rounding does not pile up at the end
        // of the list): equal run times T mean tiles_c = (T - cross * k_c) / alpha_c, hence for the CTAs c .. n-1
        //   T = (tiles left + cross * boundaries left / alpha) / sum 1 / alpha_c';  a CTA takes tiles up to the cost
        // nearest to T, at least one, and leaves one for each later CTA; the last CTA takes the rest.
        std::vector<int> begin(n + 1);
        std::vector<double> inv_tail(n + 1, 0.0);                  // sum over c' >= c of 1 / alpha_c'
        auto alpha_of = [&](int c) { return (cost->first_wave > 0 && c >= cost->first_wave) ? cost->wave2 : 1.0; };
        for (int c = n - 1; c >= 0; --c) inv_tail[c] = inv_tail[c + 1] + 1.0 / alpha_of(c);
        long long t = 0;
        for (int c = 0; c < n; ++c) {
            begin[c] = (int)t;
            const double alpha = alpha_of(c);
            const long long left = ntiles - t;
            const long long bounds_left = (ntiles - 1) / tpe - t / tpe;       // element starts strictly after tile t
            const double mean_alpha = (double)(n - c) / inv_tail[c];
            const double T = ((double)left + cost->cross * (double)bounds_left / mean_alpha) / inv_tail[c];
            double acc = 0.0;
            while (t < ntiles && (c == n - 1 || ntiles - t > (long long)(n - c - 1))) {
                const double add = alpha + ((t % tpe == 0 && t != begin[c]) ? cost->cross : 0.0);
                if (c < n - 1 && acc + 0.5 * add > T && t > begin[c]) break;
                acc += add; ++t;
            }
        }
        begin[n] = (int)ntiles;
        p.cta_tile_begin = begin;
    }
    p.el_first_cta.assign(n_el, 0);
    p.el_nparts.assign(n_el, 0);
    p.el_part_off.assign(n_el, 0);
    int c = 0, off = 0;
    for (int e = 0; e < n_el; ++e) {
        const int t0 = e * p.tiles_per_el, t1 = (e + 1) * p.tiles_per_el - 1;
        while (p.cta_tile_begin[c + 1] <= t0) ++c;
        const int first = c;
        int last = c;
        while (p.cta_tile_begin[last + 1] <= t1) ++last;
        p.el_first_cta[e] = first;
        p.el_nparts[e] = last - first + 1;
        p.el_part_off[e] = off;
        off += p.el_nparts[e];
    }
    p.total_parts = off;
}
