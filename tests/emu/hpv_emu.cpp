// TEST INFRASTRUCTURE ONLY.  Runs the CUDA kernel BODIES of hp-vpinns_b200/csrc on host threads (one
// std::thread per CUDA thread of a CTA, std::barrier for __syncthreads, CTAs one after another) so that the
// indexing / synchronisation / work-partition logic of the kernels can be exercised in the GPU-less build
// container.  It mirrors the launch sequence of hpv_api.cu with host memory.  Nothing in the product links
// or loads this file; the product has no CPU path.
#include <barrier>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "../../hp-vpinns_b200/csrc/hpv_host_prep.h"
#include "../../hp-vpinns_b200/csrc/hpv_varbwd.cuh"
#include "../../hp-vpinns_b200/csrc/hpv_points.cuh"
#include "../../hp-vpinns_b200/csrc/hpv_varbwd_tcw.cuh"     // host-side plans of the tensor-core kernels (device code is compiled out)

struct HpvEmu {
    std::barrier<>* bar;                 // the CTA barrier (__syncthreads)
    std::barrier<>** wbar;               // one barrier per warp (__syncwarp, shuffles)
    float* xchg;                         // [threads] exchange buffer of the emulated shuffles
};
void hpv_emu_barrier(HpvEmu* e) { e->bar->arrive_and_wait(); }
void hpv_emu_warp_barrier(HpvEmu* e, int warp) { e->wbar[warp]->arrive_and_wait(); }
float hpv_emu_shfl_xor(HpvEmu* e, int tid, float v, int mask) {
    e->xchg[tid] = v;
    e->wbar[tid >> 5]->arrive_and_wait();
    const float r = e->xchg[tid ^ mask];
    e->wbar[tid >> 5]->arrive_and_wait();
    return r;
}

namespace {

template <typename Body>
void run_grid(int grid, int block, size_t smem_bytes, Body body) {
    std::vector<unsigned char> smem(smem_bytes + 64);
    unsigned char* sm = smem.data();
    sm += (16 - (reinterpret_cast<uintptr_t>(sm) & 15)) & 15;
    for (int b = 0; b < grid; ++b) {
        std::barrier<> bar(block);
        const int nw = (block + 31) / 32;
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        std::vector<std::barrier<>*> wbp;
        for (int w = 0; w < nw; ++w) {
            const int n = (w + 1) * 32 <= block ? 32 : block - w * 32;
            wb.emplace_back(new std::barrier<>(n));
            wbp.push_back(wb.back().get());
        }
        std::vector<float> xchg(block, 0.0f);
        HpvEmu emu{&bar, wbp.data(), xchg.data()};
        std::vector<std::thread> th;
        th.reserve(block);
        for (int t = 0; t < block; ++t) {
            th.emplace_back([&, t]() {
                HpvCta c;
                c.tid = t; c.nthreads = block; c.bid = b; c.nblocks = grid; c.smem = sm; c.emu = &emu;
                body(c);
            });
        }
        for (auto& x : th) x.join();
    }
}

struct Key { int dim, mx, my, hp, act; };
int g_bwd_dir = 1;      // as hpv_ctx::bwd_dir: allow the directional reverse sweep where the form permits it

#define EMU_ACT(DIM, MX, MY, HP, CALL)                                             \
    if (k.act == HPV_ACT_TANH) { CALL(DIM, MX, MY, HP, HPV_ACT_TANH); }            \
    else { CALL(DIM, MX, MY, HP, HPV_ACT_SIN); }
#define EMU_MODE(HP, CALL)                                                          \
    if (k.dim == 1) {                                                               \
        if (k.mx == 0) { EMU_ACT(1, 0, 0, HP, CALL) }                               \
        else if (k.mx == 1) { EMU_ACT(1, 1, 0, HP, CALL) }                          \
        else { EMU_ACT(1, 2, 0, HP, CALL) }                                         \
    } else if (k.mx == 0 && k.my == 0) { EMU_ACT(2, 0, 0, HP, CALL) }               \
    else if (k.mx <= 1 && k.my <= 1) { EMU_ACT(2, 1, 1, HP, CALL) }                 \
    else if (k.mx == 2 && k.my <= 1) { EMU_ACT(2, 2, 1, HP, CALL) }                 \
    else { EMU_ACT(2, 2, 2, HP, CALL) }
#define EMU_DISPATCH(CALL)                                                          \
    if (k.hp == 8) { EMU_MODE(8, CALL) }                                            \
    else if (k.hp == 20) { EMU_MODE(20, CALL) }                                     \
    else if (k.hp == 32) { EMU_MODE(32, CALL) }                                     \
    else return -4;

int emu_fwd(const Key& k, const HpvVarArgs& a, int grid, size_t smem) {
#define CALL(DIM, MX, MY, HP, ACT) run_grid(grid, HPV_THREADS, smem, [&](const HpvCta& c) { hpv_varfwd_body<DIM, MX, MY, HP, ACT>(c, a); })
    EMU_DISPATCH(CALL)
#undef CALL
    return 0;
}

int emu_bwd(const Key& k, const HpvBwdArgs& a, int grid, int block) {
#define CALL(DIM, MX, MY, HP, ACT)                                                                       \
    {                                                                                                    \
        HpvBwdSmem<DIM, MX, MY, HP> L(a.v.nhid, block);                                 \
        run_grid(grid, block, (size_t)L.total * 4, [&](const HpvCta& c) { hpv_mlpbwd_body<DIM, MX, MY, HP, ACT>(c, a); }); \
    }
    if (k.dim == 2 && k.mx == 1 && k.my == 1 && g_bwd_dir && a.pts == nullptr) {
        bool dir = true;
        for (int t = 0; t < a.v.n_terms; ++t)
            for (int f = 0; f < HPV_NFIELDS; ++f) if (a.v.terms[t].a1[f] != 0.0f) dir = false;
        if (dir) {
            if (k.hp == 8) { EMU_ACT(2, 1, 0, 8, CALL) }
            else if (k.hp == 20) { EMU_ACT(2, 1, 0, 20, CALL) }
            else if (k.hp == 32) { EMU_ACT(2, 1, 0, 32, CALL) }
            else return -4;
            return 0;
        }
    }
    EMU_DISPATCH(CALL)
#undef CALL
    return 0;
}

int emu_pts(const Key& k, const HpvPointArgs& a, float* gbar, int grid) {
    const size_t smem = (size_t)(HPV_THREADS + hpv_slot_floats(k.dim, k.mx, k.my, k.hp, HPV_THREADS)) * 4;
#define CALL(DIM, MX, MY, HP, ACT) run_grid(grid, HPV_THREADS, smem, [&](const HpvCta& c) { hpv_points_body<DIM, MX, MY, HP, ACT>(c, a, gbar); })
    EMU_DISPATCH(CALL)
#undef CALL
    return 0;
}

void reduce_grad(const std::vector<float>& part, int n_parts, int stride, int n, std::vector<float>& out, int accumulate) {
    HpvGradReduceArgs g;
    g.grad_part = part.data(); g.n_parts = n_parts; g.stride = stride; g.n = n; g.grad_pad = out.data(); g.accumulate = accumulate;
    run_grid((n + 31) / 32, 256, 8 * 32 * 4, [&](const HpvCta& c) { hpv_gradreduce_body(c, g); });
}

}  // namespace

extern "C" {

void hpv_emu_set_bwd_dir(int on) { g_bwd_dir = on; }

// Host-side planning helpers exposed for property tests.
// hpv_partition: returns n_ctas; cta_tile_begin [n_ctas+1], el_first_cta / el_nparts / el_part_off [n_el].
int hpv_emu_partition(int n_el, int pts_per_el, int tile_pts, int max_ctas, int cta_pts, int* tiles_per_el, int* total_parts,
                      int* cta_tile_begin, int* el_first_cta, int* el_nparts, int* el_part_off) {
    HpvPartition p;
    hpv_partition(p, n_el, pts_per_el, tile_pts, max_ctas, cta_pts);
    *tiles_per_el = p.tiles_per_el; *total_parts = p.total_parts;
    for (int c = 0; c <= p.n_ctas; ++c) cta_tile_begin[c] = p.cta_tile_begin[c];
    for (int e = 0; e < n_el; ++e) { el_first_cta[e] = p.el_first_cta[e]; el_nparts[e] = p.el_nparts[e]; el_part_off[e] = p.el_part_off[e]; }
    return p.n_ctas;
}

// the same with the cost model of the tensor-core forward kernel's launch plan
int hpv_emu_partition_weighted(int n_el, int pts_per_el, int tile_pts, int max_ctas, double cross, int first_wave, double wave2,
                               int* tiles_per_el, int* cta_tile_begin) {
    HpvPartition p;
    HpvPartitionCost pc; pc.cross = cross; pc.first_wave = first_wave; pc.wave2 = wave2;
    hpv_partition(p, n_el, pts_per_el, tile_pts, max_ctas, 0, &pc);
    *tiles_per_el = p.tiles_per_el;
    for (int c = 0; c <= p.n_ctas; ++c) cta_tile_begin[c] = p.cta_tile_begin[c];
    return p.n_ctas;
}

// padded parameter index -> compact gradient index of the reverse sweep (or -1), and the compact length
int hpv_emu_gw_of_padded(int dim, int hp, int nhid, int ip) { return hpv_gw_of_padded(dim, hp, nhid, ip); }
int hpv_emu_gw_n(int dim, int hp, int nhid) { return hpv_gw_n(dim, hp, nhid); }
int hpv_emu_theta_pad_n(int dim, int hp, int nhid) { return hpv_theta_pad_n(dim, hp, nhid); }
int hpv_emu_off_wt(int dim, int hp, int l) { return hpv_off_wt(dim, hp, l); }
int hpv_emu_off_wl(int dim, int hp, int l) { return hpv_off_wl(dim, hp, l); }

// Plans of the tensor-core kernels: TMEM columns and shared-memory layouts (offsets in floats).
// out[0..7] forward: total, B, bar, th, part, tab0, P, G;  out[8..13] reverse sweep (FFMA weight gradients): total, B, bar,
// slots, gw, slot_sz;  out[14..18] reverse sweep (tensor-core weight gradients): total, B, bar, op, slots
int hpv_emu_tc_plan(int dim, int hp, int nch, int nch1, int nhid, int Q, int rows, int n_terms, int ltab, int rtab, int* out) {
    HpvVarArgs a;
    memset(&a, 0, sizeof(a));
    a.Q = Q; a.rows = rows; a.n_terms = n_terms; a.nhid = nhid;
    for (int t = 0; t < n_terms; ++t) { a.terms[t].ltab = t == 0 ? ltab : rtab; a.terms[t].rtab = t == 0 ? rtab : ltab; }
    const HpvFwdTcSmem f = hpv_fwd_tc_smem(a, dim, hp, nch);
    out[0] = f.total; out[1] = f.B; out[2] = f.bar; out[3] = f.th; out[4] = f.part; out[5] = f.tab[ltab]; out[6] = f.P; out[7] = f.G;
    const HpvBwdTcSmem b = hpv_bwd_tc_smem(dim, hp, nch, nch1, nhid);
    out[8] = b.total; out[9] = b.B; out[10] = b.bar; out[11] = b.slots; out[12] = b.gw; out[13] = b.slot_sz;
    const HpvBwdTcwSmem w = hpv_bwd_tcw_smem(dim, hp, nch, nhid);
    out[14] = w.total; out[15] = w.B; out[16] = w.bar; out[17] = w.op; out[18] = w.slots;
    return hpv_tc_tmem_need(nch, hp);
}
int hpv_emu_tcw_tmem_need(int nch, int hp, int nhid) { return hpv_tcw_tmem_need(nch, hp, nhid); }
int hpv_emu_tcw_supported(int nch, int hp, int nhid) { return hpv_tcw_supported(nch, hp, nhid) ? 1 : 0; }

// Variational loss forward (+ backward when grad_theta != NULL) on emulated CTAs.
//   n_ctas_fwd / n_ctas_bwd / bwd_block choose the launch geometry (to exercise the split-element paths).
int hpv_emu_varloss(int dim, const int* layers, int n_layers, int act, int Q, const double* xi, const double* w,
                    int N, const double* T, const double* D1, const double* D2, const double* d1b, int problem,
                    int var_form, double V, int n_el, const double* lo, const double* hi, const int* ntest, int ntx,
                    int nty, const double* F_ext, const double* theta, double eps, int n_ctas_fwd, int n_ctas_bwd,
                    int bwd_block, double* loss_out, float* res_out, double* el_loss_out, double* grad_theta,
                    double* grad_eps) {
    std::string err;
    HpvNet net;
    if (!hpv_net_setup(net, dim, layers, n_layers, act, err)) return -1;
    HpvForm fm;
    if (!hpv_form_setup(fm, problem, var_form, V, err)) return -1;
    if (dim == 1) nty = 1;
    std::vector<float> tabs[HPV_NTAB];
    hpv_build_tables(Q, N, w, T, D1, D2, d1b, fm.fold_boundary, tabs);
    std::vector<float> nat[HPV_NTAB];
    hpv_natural_tables(Q, tabs, nat);
    std::vector<float> xi1(Q);
    for (int q = 0; q < Q; ++q) xi1[q] = (float)(xi[q] + 1.0);
    std::vector<float> th;
    hpv_pad_theta(net, theta, th);
    float epsf = (float)eps;
    std::vector<float> geom((size_t)n_el * 4);
    std::vector<int> nt((size_t)n_el * 2);
    for (int e = 0; e < n_el; ++e) {
        geom[4 * e] = (float)lo[e * dim]; geom[4 * e + 1] = (float)((hi[e * dim] - lo[e * dim]) / 2);
        geom[4 * e + 2] = 0; geom[4 * e + 3] = 1;
        if (dim == 2) { geom[4 * e + 2] = (float)lo[e * dim + 1]; geom[4 * e + 3] = (float)((hi[e * dim + 1] - lo[e * dim + 1]) / 2); }
        nt[2 * e] = ntest ? ntest[e * dim] : ntx;
        nt[2 * e + 1] = (dim == 2) ? (ntest ? ntest[e * dim + 1] : nty) : 1;
    }
    const size_t nf = (size_t)n_el * nty * ntx;
    std::vector<float> F(nf, 0.0f), Res(nf, 0.0f), el_loss(n_el, 0.0f);
    if (F_ext) for (size_t i = 0; i < nf; ++i) F[i] = (float)F_ext[i];
    const int rows = (dim == 2) ? Q : 1;
    HpvPartition part;
    hpv_partition(part, n_el, rows * Q, HPV_FWD_TILE, n_ctas_fwd);          // no minimum: the tests want elements split finely
    std::vector<float> Upart((size_t)part.total_parts * HPV_NP * HPV_NP, 0.0f);
    std::vector<unsigned int> counters(n_el + 1, 0u);
    double loss = 0.0;

    HpvVarArgs a;
    memset(&a, 0, sizeof(a));
    a.theta_pad = th.data(); a.theta_pad_n = net.theta_pad_n; a.nhid = net.nhid; a.eps = &epsf;
    a.off_wo = hpv_off_wo(net.dim, net.hp, net.nhid);
    a.Q = Q; a.rows = rows; a.xi1 = xi1.data();
    for (int t = 0; t < HPV_NTAB; ++t) { a.tab[t] = tabs[t].data(); a.tabN[t] = nat[t].data(); }
    a.QP = hpv_align4(Q);
    a.n_el = n_el; a.el_geom = geom.data(); a.el_ntest = nt.data(); a.ntx = ntx; a.nty = nty;
    a.F = fm.has_rhs && F_ext ? F.data() : nullptr;
    a.n_terms = fm.n_terms;
    for (int t = 0; t < HPV_MAX_TERMS; ++t) a.terms[t] = fm.terms[t];
    a.tile_pts = HPV_FWD_TILE; a.tiles_per_el = part.tiles_per_el; a.n_ctas = part.n_ctas;
    a.cta_tile_begin = part.cta_tile_begin.data(); a.el_first_cta = part.el_first_cta.data();
    a.el_part_off = part.el_part_off.data(); a.el_nparts = part.el_nparts.data();
    a.Upart = Upart.data(); a.el_done = counters.data(); a.n_done = counters.data() + n_el;
    a.Res = Res.data(); a.el_loss = el_loss.data(); a.loss = &loss; a.loss_scale = 1.0f;

    Key k{dim, fm.mx, fm.my, net.hp, act};
    int cmx = fm.mx, cmy = fm.my; hpv_canon_mode(dim, cmx, cmy);
    const HpvFwdSmem fs = hpv_fwd_smem(a, hpv_slot_floats(dim, cmx, cmy, net.hp, HPV_THREADS));
    int r = emu_fwd(k, a, part.n_ctas, (size_t)fs.total * 4);
    if (r) return r;
    if (counters[n_el] != 0u) return -7;                 // self-resetting counters must be back at zero
    for (int e = 0; e < n_el; ++e) if (counters[e] != 0u) return -7;
    if (loss_out) *loss_out = loss;
    if (res_out) memcpy(res_out, Res.data(), nf * sizeof(float));
    if (el_loss_out) for (int e = 0; e < n_el; ++e) el_loss_out[e] = el_loss[e];
    if (!grad_theta) return 0;

    // backward: K1 adjoint projection, K2 MLP reverse sweep, K3 reduction
    const int npts = n_el * rows * Q;
    std::vector<float> Gbar((size_t)fm.n_terms * npts, 0.0f);
    HpvAdjArgs aa;
    aa.v = a; aa.Gbar = Gbar.data(); aa.slabs_per_el = (rows + HPV_ADJ_RS - 1) / HPV_ADJ_RS;
    run_grid(n_el * aa.slabs_per_el, HPV_THREADS, (size_t)hpv_adj_smem(a).total * 4,
             [&](const HpvCta& c) { hpv_adjproj_body(c, aa); });
    const int stride = hpv_align4(net.theta_pad_n + 1);
    int ntiles = (npts + bwd_block - 1) / bwd_block;
    int grid_b = n_ctas_bwd < ntiles ? n_ctas_bwd : ntiles;
    std::vector<float> gpart((size_t)grid_b * stride, 0.0f), gpad(stride, 0.0f);
    HpvBwdArgs ba;
    ba.v = a; ba.v.grad_part = gpart.data(); ba.v.grad_stride = stride;
    ba.Gbar = Gbar.data(); ba.n_points = npts; ba.pts = nullptr; ba.stagger_ns = 0;
    r = emu_bwd(k, ba, grid_b, bwd_block);
    if (r) return r;
    reduce_grad(gpart, grid_b, stride, net.theta_pad_n + 1, gpad, 0);
    for (int i = 0; i < net.n_theta; ++i) grad_theta[i] = gpad[net.pad_index[i]];
    if (grad_eps) *grad_eps = gpad[net.theta_pad_n];
    return 0;
}

// Point evaluation / point loss (+ gradient) on emulated CTAs.
int hpv_emu_points(int dim, const int* layers, int n_layers, int act, const double* theta, double eps, int n,
                   const double* pts, const double* target, const double* a0, const double* a1, double weight,
                   int mode, double* u, double* d1, double* d2, double* loss_out, double* grad_theta, double* grad_eps,
                   int bwd_block) {
    std::string err;
    HpvNet net;
    if (!hpv_net_setup(net, dim, layers, n_layers, act, err)) return -1;
    std::vector<float> th;
    hpv_pad_theta(net, theta, th);
    float epsf = (float)eps;
    std::vector<float> p((size_t)n * dim), tg(n, 0.0f), ou(n), od1((size_t)n * dim), od2((size_t)n * dim), resid(n), gbar(n);
    for (size_t i = 0; i < p.size(); ++i) p[i] = (float)pts[i];
    HpvPointArgs a;
    memset(&a, 0, sizeof(a));
    a.theta_pad = th.data(); a.theta_pad_n = net.theta_pad_n; a.nhid = net.nhid; a.eps = &epsf;
    a.off_wo = hpv_off_wo(net.dim, net.hp, net.nhid);
    a.n = n; a.pts = p.data(); a.out_u = ou.data(); a.out_d1 = od1.data(); a.out_d2 = od2.data();
    int mx = mode, my = mode;
    if (target) {
        for (int i = 0; i < n; ++i) tg[i] = (float)target[i];
        for (int f = 0; f < HPV_NFIELDS; ++f) { a.a0[f] = (float)a0[f]; a.a1[f] = a1 ? (float)a1[f] : 0.0f; }
        a.target = tg.data(); a.weight = (float)weight; a.resid = resid.data();
        hpv_mode_of_coef(dim, a.a0, a.a1, mx, my);
    }
    hpv_canon_mode(dim, mx, my);
    int grid = (n + HPV_THREADS - 1) / HPV_THREADS;
    std::vector<float> blk(grid, 0.0f);
    a.blk_loss = blk.data(); a.n_ctas = grid;
    Key k{dim, mx, my, net.hp, act};
    int r = emu_pts(k, a, target ? gbar.data() : nullptr, grid);
    if (r) return r;
    if (u) for (int i = 0; i < n; ++i) u[i] = ou[i];
    if (d1) for (size_t i = 0; i < od1.size(); ++i) d1[i] = od1[i];
    if (d2) for (size_t i = 0; i < od2.size(); ++i) d2[i] = od2[i];
    if (target && loss_out) { double s = 0; for (float v : blk) s += v; *loss_out = s; }
    if (!target || !grad_theta) return 0;
    const int stride = hpv_align4(net.theta_pad_n + 1);
    int ntiles = (n + bwd_block - 1) / bwd_block;
    int grid_b = ntiles < 3 ? ntiles : 3;
    std::vector<float> gpart((size_t)grid_b * stride, 0.0f), gpad(stride, 0.0f);
    HpvBwdArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.v.theta_pad = th.data(); ba.v.theta_pad_n = net.theta_pad_n; ba.v.nhid = net.nhid; ba.v.eps = &epsf;
    ba.v.off_wo = hpv_off_wo(net.dim, net.hp, net.nhid);
    ba.v.Q = 1; ba.v.rows = 1; ba.v.n_terms = 1; ba.v.terms[0] = hpv_term_zero();
    for (int f = 0; f < HPV_NFIELDS; ++f) { ba.v.terms[0].a0[f] = a.a0[f]; ba.v.terms[0].a1[f] = a.a1[f]; }
    ba.v.grad_part = gpart.data(); ba.v.grad_stride = stride;
    ba.Gbar = gbar.data(); ba.n_points = n; ba.pts = p.data();
    r = emu_bwd(k, ba, grid_b, bwd_block);
    if (r) return r;
    reduce_grad(gpart, grid_b, stride, net.theta_pad_n + 1, gpad, 0);
    for (int i = 0; i < net.n_theta; ++i) grad_theta[i] = gpad[net.pad_index[i]];
    if (grad_eps) *grad_eps = gpad[net.theta_pad_n];
    return 0;
}

}  // extern "C"
