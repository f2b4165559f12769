// libhpv C ABI (include/hpv.h): context, device memory, launch sequencing.  No arithmetic of the hot path
// lives here -- that is all in the kernels -- and there is no CPU fallback: without a CUDA device every
// compute entry point fails with HPV_ERR_CUDA.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/hpv.h"
#include "hpv_host_prep.h"
#include "hpv_launch.h"

namespace {

std::string g_create_error;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) n = count ? count : 1; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// Scope guard for the temporary device buffers of a single call: released on every return path (HPV_CK returns early).
template <typename T>
struct ScopedBuf : DevBuf<T> {
    ScopedBuf() = default;
    ScopedBuf(const ScopedBuf&) = delete;
    ScopedBuf& operator=(const ScopedBuf&) = delete;
    ~ScopedBuf() { this->release(); }
};

struct PointSet {
    bool active = false;
    int n = 0, mx = 0, my = 0, n_ctas = 0;
    float a0[HPV_NFIELDS], a1[HPV_NFIELDS];
    float weight = 1.0f;
    DevBuf<float> pts, target, resid, gbar, blk_loss;
};

}  // namespace

struct hpv_ctx {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    std::string err;
    long long launches = 0;

    HpvNet net; bool have_net = false;
    int Q = 0; std::vector<double> xi, w; bool have_quad = false;
    int N = 0; std::vector<double> T, D1, D2, d1b; bool have_tabs = false, have_d1b = false;
    int problem = -1, var_form = -1; double V = 1.0; HpvForm form; bool have_form = false;
    int n_el = 0, ntx = 0, nty = 0; bool have_el = false, has_F = false;
    bool ready = false;
    bool grad_ready = false;           // network-dependent reduction buffers (grad_part, redbuf) match the current network

    // parameters / optimiser
    DevBuf<float> theta_pad, eps;
    DevBuf<double> master, adam_m, adam_v, grad_out;
    DevBuf<int> pad_index, pad_index2, ref_index;
    // gradient exchange over peer memory (hpv_peer_export / hpv_peer_connect)
    DevBuf<float> peer_inbox_own;
    DevBuf<unsigned> peer_flags_own, peer_err;
    int peer_n = 0, peer_rank = 0, peer_nvp = 0, peer_nchunks = 0;
    unsigned peer_seq = 0;
    unsigned long long peer_timeout_ns = 20000000000ull;     // HPV_PEER_TIMEOUT_S: how long a rank waits for its peers
    float* peer_inbox[HPV_MAX_PEERS] = {nullptr};
    unsigned* peer_flags[HPV_MAX_PEERS] = {nullptr};
    std::vector<void*> peer_opened;
    DevBuf<double> adam_clock;         // two buffers of {beta1^t, beta2^t, t}; the update reads [adam_parity], writes the other
    int adam_parity = 0;
    // quadrature / tables
    DevBuf<float> xi1, tab[HPV_NTAB], tabN[HPV_NTAB];
    // elements
    DevBuf<float> geom, F, Res, el_loss, Upart, Gbar;
    DevBuf<int> ntest, cta_tile_begin, el_first_cta, el_part_off, el_nparts;
    DevBuf<unsigned int> counters;                 // [n_el] el_done, then n_done
    DevBuf<double> loss;
    HpvPartition part;
    // backward
    DevBuf<float> grad_part, redbuf;
    int bwd_block = 0, bwd_grid = 0, bwd_ctas_per_sm = 0, grad_stride = 0, loss_off = 0;
    size_t bwd_smem = 0, fwd_smem = 0, adj_smem = 0;
    bool bwd_dir = true;               // allow the directional reverse sweep (HPV_BWD_DIR=0 disables)
    int bwd_stagger_ns = 0;            // start offset between the warp rows of the reverse sweep (HPV_BWD_STAGGER_NS)
    bool adam_direct = true;           // Adam writes the constant-memory mirrors in place (HPV_ADAM_DIRECT=0 disables)
    // training steps replayed from a CUDA graph (hpv_train_steps): two consecutive steps are captured once per
    // configuration (the optimizer clock is double-buffered, so the launch arguments have period 2) and re-launched
    bool use_graph = true;             // HPV_GRAPH=0: plain launches
    cudaGraphExec_t step_graph[2] = {nullptr, nullptr};   // one step each, for the two parities of the optimizer clock
    unsigned long long epoch = 1, graph_epoch = 0;     // every call that changes a launch argument bumps `epoch`
    unsigned long long plain_epoch = 0; int plain_steps = 0;   // plain steps run since the last change (lazy set-up done?)
    int graph_launches = 0; unsigned graph_kinds = 0, used_kinds = 0; bool graph_hist = false, capturing = false;
    bool last_wrote[3] = {false, false, false}, graph_wrote[3] = {false, false, false};   // mirrors the (captured) Adam launch writes
    DevBuf<float> hist_ring; DevBuf<double> hist_t0;   // device-side loss history of graph-replayed steps
    double adam_t = 0.0;                               // host copy of the optimizer's step counter
    bool fwd_tc = true;                // forward kernel in its tensor-core form (HPV_FWD_TC=0: the FP32-FFMA form)
    bool fwd_tc_active = false;        // ... and the current network / form / rule admit it (decided by ensure_ready)
    // Reverse sweep in its tensor-core form (hpv_varbwd_tc.cuh): -1 = automatic (small batches only: below ~2 warp
    // tiles per SM the FFMA sweep cannot fill the machine and the tensor-core form is 2x faster, C2: 23 vs 47 us; on
    // large batches the FFMA sweep is still the faster one, C3: 133 vs 152 us), HPV_BWD_TC=0/1 forces it off/on
    int bwd_tc = -1;
    bool bwd_tc_active = false;
    bool bwd_tcw = false;              // ... with the weight gradients on the tensor cores too (HPV_BWD_TCW=1; measured slower and
                                       // less accurate than the FMA-pipe weight gradients: DESIGN.md, kept for A/B measurements)
    bool bwd_tcw_active = false;
    int fwd_ctas_per_sm = 0, adj_grid = 0, slabs_per_el = 0, part_tile = HPV_FWD_TILE;
    PointSet ps[HPV_MAX_POINT_SETS];
    // training configuration
    double wv = 1.0; unsigned mask = 0; int train_eps = 0;
    double lr = 1e-3, b1 = 0.9, b2 = 0.999, eps_hat = 1e-8;
    std::vector<float> host_f32;
    std::vector<double> host_f64;
    // pinned staging of hpv_set_params (no host synchronisation: the copies are asynchronous and the buffer is
    // only rewritten after `param_staged` says the previous upload has been consumed)
    DevBuf<unsigned char> param_blob;                 // device side of the staged upload (scattered by a kernel)
    unsigned char* param_stage = nullptr; size_t param_stage_bytes = 0;
    cudaEvent_t param_staged = nullptr; bool param_stage_busy = false;
    // constant-memory mirrors of theta_pad, one per kernel translation unit kind (forward, reverse sweep, points)
    float* mirror[3] = {nullptr, nullptr, nullptr};
    bool mirror_stale[3] = {true, true, true};
};

namespace {

// Which context's parameters each constant-memory copy (device x padded width x kernel kind) holds right now.
std::mutex g_owner_mutex;
hpv_ctx* g_owner[16][3][3];

int fail(hpv_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define HPV_CK(call)                                                                                 \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(c, HPV_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

template <typename T>
int upload(hpv_ctx* c, DevBuf<T>& b, const std::vector<T>& h) {
    HPV_CK(b.alloc(h.size()));
    if (!h.empty()) HPV_CK(cudaMemcpyAsync(b.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));      // the host vector may be a temporary
    return HPV_OK;
}

HpvKernelKey key_of(const hpv_ctx* c, int mx, int my) {
    HpvKernelKey k;
    k.dim = c->net.dim; k.mx = mx; k.my = my; k.hp = c->net.hp; k.act = c->net.act; k.dir = 0; k.wg = 0;
    hpv_canon_mode(k.dim, k.mx, k.my);
    return k;
}

// Key of the reverse sweep of the variational loss: the directional mode when the form allows it
// (hpv_form_directional); HPV_BWD_DIR=0 in the environment keeps the two-tangent sweep (A/B measurements).
HpvKernelKey bwd_key_of(const hpv_ctx* c) {
    HpvKernelKey k = key_of(c, c->form.mx, c->form.my);
    k.dir = (c->bwd_dir && hpv_form_directional(c->net.dim, c->form)) ? 1 : 0;
    return k;
}

// Bring the constant-memory copy of the parameters that kernels of `kind` read up to date (stream-ordered).
int refresh_mirror(hpv_ctx* c, int kind) {
    c->used_kinds |= 1u << kind;
    if (!c->mirror[kind]) {
        HpvLaunch l; memset(&l, 0, sizeof(l));
        long long out = 0;
        l.kind = kind; l.op = 3; l.out = &out;
        HPV_CK(hpv_dispatch(key_of(c, 0, 0), l));
        c->mirror[kind] = reinterpret_cast<float*>((uintptr_t)out);
        c->mirror_stale[kind] = true;
    }
    std::lock_guard<std::mutex> lock(g_owner_mutex);
    const int hpi = c->net.hp == 8 ? 0 : (c->net.hp == 20 ? 1 : 2);
    hpv_ctx*& owner = g_owner[c->device & 15][hpi][kind];
    if (owner != c) {
        // another context's kernels may still be reading this copy: let them finish before overwriting it
        if (owner && owner->stream != c->stream) HPV_CK(cudaStreamSynchronize(owner->stream));
        owner = c;
        c->mirror_stale[kind] = true;
    }
    if (c->mirror_stale[kind]) {
        HPV_CK(cudaMemcpyAsync(c->mirror[kind], c->theta_pad.p, c->net.theta_pad_n * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        c->mirror_stale[kind] = false;
    }
    return HPV_OK;
}

void theta_changed(hpv_ctx* c) { c->mirror_stale[0] = c->mirror_stale[1] = c->mirror_stale[2] = true; }

void fill_var_args(hpv_ctx* c, HpvVarArgs& a) {
    memset(&a, 0, sizeof(a));
    a.theta_pad = c->theta_pad.p; a.theta_pad_n = c->net.theta_pad_n; a.nhid = c->net.nhid; a.eps = c->eps.p;
    a.off_wo = hpv_off_wo(c->net.dim, c->net.hp, c->net.nhid);

    a.Q = c->Q; a.rows = (c->net.dim == 2) ? c->Q : 1; a.xi1 = c->xi1.p;
    for (int t = 0; t < HPV_NTAB; ++t) { a.tab[t] = c->tab[t].p; a.tabN[t] = c->tabN[t].p; }
    a.QP = hpv_align4(c->Q);
    a.n_el = c->n_el; a.el_geom = c->geom.p; a.el_ntest = c->ntest.p; a.ntx = c->ntx; a.nty = c->nty;
    a.F = c->has_F ? c->F.p : nullptr;
    a.n_terms = c->form.n_terms;
    for (int t = 0; t < HPV_MAX_TERMS; ++t) a.terms[t] = c->form.terms[t];
    a.tile_pts = c->part_tile; a.tiles_per_el = c->part.tiles_per_el; a.n_ctas = c->part.n_ctas;
    a.cta_tile_begin = c->cta_tile_begin.p; a.el_first_cta = c->el_first_cta.p;
    a.el_part_off = c->el_part_off.p; a.el_nparts = c->el_nparts.p;
    a.Upart = c->Upart.p; a.el_done = c->counters.p; a.n_done = c->counters.p + c->n_el;
    a.Res = c->Res.p; a.el_loss = c->el_loss.p; a.loss = c->loss.p;
    a.grad_part = c->grad_part.p; a.grad_stride = c->grad_stride; a.grad_pad = c->redbuf.p;
    a.bwd_done = nullptr; a.loss_scale = (float)c->wv;
}

// Launch plan of the MLP reverse sweep for `n_points` points: one CTA per SM of W warps, every warp running the
// same number of 32-point tiles (hpv_mlpbwd_body).  W is bounded by the kernel's launch bounds (registers) and by
// the shared-memory plan; among the admissible W the plan takes the fewest tiles per warp, and for that count the
// fewest warps: at C3 (12 800 tiles on 148 SMs) 11 warps x 8 tiles instead of 12 x 8 -- the same makespan with
// less contention and no empty tile slots -- and a handful of boundary points spread as single-warp CTAs.
int plan_bwd(hpv_ctx* c, const HpvKernelKey& k, long long n_points, int& block, int& grid, size_t& smem) {
    HpvVarArgs va; memset(&va, 0, sizeof(va));
    va.theta_pad_n = c->net.theta_pad_n; va.nhid = c->net.nhid; va.off_wo = hpv_off_wo(c->net.dim, c->net.hp, c->net.nhid);
    HpvBwdArgs ba; memset(&ba, 0, sizeof(ba)); ba.v = va;
    HpvLaunch l; memset(&l, 0, sizeof(l));
    long long out = 0;
    l.kind = HPV_K_MLPBWD; l.bwd = &ba; l.out = &out;
    l.op = 4;
    HPV_CK(hpv_dispatch(k, l));
    int wmax = (int)out / 32;
    const long long n_wt = (n_points + 31) / 32;
    int best_w = 0;
    long long best_iters = 0;
    size_t best_smem = 0;
    int forced = 0;
    if (const char* ev = getenv("HPV_BWD_WARPS")) forced = atoi(ev);           // tuning override
    for (int w = wmax; w >= 1; --w) {
        l.op = 2; l.block = 32 * w;
        HPV_CK(hpv_dispatch(k, l));
        const size_t sm = (size_t)out;
        if (sm > 227 * 1024) continue;
        // leave >= 12 KB of the SM's 228 KB to the L1 (register spills, Gbar / geometry reads): at C4 15 warps with
        // 215 KB run 1 953 us, 16 warps with 229 KB 1 997 us (profiles/r02r)
        if (sm > 216 * 1024 && w > 1 && !forced) continue;
        const long long iters = (n_wt + (long long)c->n_sm * w - 1) / ((long long)c->n_sm * w);
        if (forced ? (w == forced || !best_w) : (!best_w || iters <= best_iters)) { best_w = w; best_iters = iters; best_smem = sm; }
        if (forced && w == forced) break;
    }
    if (!best_w) return fail(c, HPV_ERR_LIMIT, "network too deep/wide for the shared-memory plan of the MLP reverse sweep");
    // Shape of the W warps of an SM: one CTA.  (With the earlier 168-register build, 4-warp CTAs were faster than
    // one 12-warp CTA -- the warps of a big CTA start in lockstep and convoy through the phases --; with 128 registers
    // and 15 warps the single CTA wins at C3 and C4, and several small CTAs per SM interact badly with the
    // programmatic dependent launches (profiles/r02q, r02r).  HPV_BWD_CTA_WARPS regroups for experiments.)
    int wcta = best_w;
    int force_cta = 0;
    if (const char* ev = getenv("HPV_BWD_CTA_WARPS")) force_cta = atoi(ev);
    if (force_cta > 0 && best_w % force_cta == 0) {
        l.op = 2; l.block = 32 * force_cta;
        HPV_CK(hpv_dispatch(k, l));
        wcta = force_cta; best_smem = (size_t)out;
    }
    block = 32 * wcta; smem = best_smem;
    const long long n_grp = (n_wt + wcta - 1) / wcta;
    const long long slots = (long long)c->n_sm * (best_w / wcta);
    grid = (int)(n_grp < slots ? (n_grp < 1 ? 1 : n_grp) : slots);
    return HPV_OK;
}

// Buffers of the gradient reduction.  They depend on the network only (padded parameter count), so that a step
// made of point-wise losses alone (wv = 0) needs no element batch; hpv_set_network invalidates them.
int ensure_grad_buffers(hpv_ctx* c, int min_parts) {
    if (!c->grad_ready) {
        c->grad_stride = hpv_align4(c->net.theta_pad_n + 1);
        c->loss_off = ((c->net.theta_pad_n + 1 + 31) / 32) * 32;       // the losses start a 32-entry chunk of their own
        HPV_CK(c->redbuf.alloc(c->loss_off + 32));
        HPV_CK(cudaMemsetAsync(c->redbuf.p, 0, (c->loss_off + 32) * sizeof(float), c->stream));
        c->grad_part.release();
    }
    int parts = min_parts;
    if (parts < c->n_sm * 4) parts = c->n_sm * 4;                      // point-loss launches reuse the buffer
    if (!c->grad_part.p || c->grad_part.n < (size_t)parts * c->grad_stride) HPV_CK(c->grad_part.alloc((size_t)parts * c->grad_stride));
    c->grad_ready = true;
    return HPV_OK;
}

int ensure_ready(hpv_ctx* c) {
    if (c->ready) return HPV_OK;
    if (!c->have_net) return fail(c, HPV_ERR_STATE, "hpv_set_network has not been called");
    if (!c->have_quad) return fail(c, HPV_ERR_STATE, "hpv_set_quadrature has not been called");
    if (!c->have_tabs) return fail(c, HPV_ERR_STATE, "hpv_set_test_tables has not been called");
    if (!c->have_form) return fail(c, HPV_ERR_STATE, "hpv_set_form has not been called");
    if (!c->have_el) return fail(c, HPV_ERR_STATE, "hpv_set_elements has not been called");
    const int pdim = (c->problem == HPV_POISSON1D) ? 1 : 2;
    if (pdim != c->net.dim) return fail(c, HPV_ERR_ARG, "network input dimension does not match the problem");
    if (c->ntx > c->N || (pdim == 2 && c->nty > c->N)) return fail(c, HPV_ERR_ARG, "more test functions requested than the tables hold");
    if (c->form.fold_boundary) {
        if (!c->have_d1b) return fail(c, HPV_ERR_STATE, "Poisson-1D var_form 3 needs d1_bound in hpv_set_test_tables");
        if (fabs(c->xi[0] + 1.0) > 1e-12 || fabs(c->xi[c->Q - 1] - 1.0) > 1e-12)
            return fail(c, HPV_ERR_ARG, "Poisson-1D var_form 3 needs Gauss-Lobatto nodes (xi[0] = -1, xi[Q-1] = 1)");
    }
    std::vector<float> tabs[HPV_NTAB];
    hpv_build_tables(c->Q, c->N, c->w.data(), c->T.data(), c->D1.data(), c->D2.data(),
                     c->have_d1b ? c->d1b.data() : nullptr, c->form.fold_boundary, tabs);
    std::vector<float> nat[HPV_NTAB];
    hpv_natural_tables(c->Q, tabs, nat);
    for (int t = 0; t < HPV_NTAB; ++t) {
        int r = upload(c, c->tab[t], tabs[t]); if (r) return r;
        r = upload(c, c->tabN[t], nat[t]); if (r) return r;
    }
    std::vector<float> xi1(c->Q);
    for (int q = 0; q < c->Q; ++q) xi1[q] = (float)(c->xi[q] + 1.0);
    { int r = upload(c, c->xi1, xi1); if (r) return r; }

    // forward launch plan: the tensor-core form (hpv_varfwd_tc.cuh) when its TMEM / shared-memory plan fits, else FFMA
    HpvVarArgs a; fill_var_args(c, a);
    const HpvKernelKey k = key_of(c, c->form.mx, c->form.my);
    const int nch = hpv_mode_nch(k.dim, k.mx, k.my);
    c->fwd_tc_active = false;
    if (c->fwd_tc && hpv_tc_supported(nch, k.hp)) {
        const size_t sm_tc = (size_t)hpv_fwd_tc_smem(a, k.dim, k.hp, nch).total * 4;
        if (sm_tc <= 227 * 1024) {
            HpvLaunch l; memset(&l, 0, sizeof(l));
            long long out = 0;
            l.kind = HPV_K_VARFWD_TC; l.op = 1; l.block = HPV_THREADS; l.smem = sm_tc; l.out = &out;
            HPV_CK(hpv_dispatch(k, l));
            if (out >= 1) { c->fwd_tc_active = true; c->fwd_smem = sm_tc; c->fwd_ctas_per_sm = (int)out; }
        }
    }
    if (!c->fwd_tc_active) {
        const HpvFwdSmem fs = hpv_fwd_smem(a, hpv_slot_floats(k.dim, k.mx, k.my, k.hp, HPV_THREADS));
        c->fwd_smem = (size_t)fs.total * 4;
        if (c->fwd_smem > 227 * 1024) return fail(c, HPV_ERR_LIMIT, "forward kernel shared-memory plan exceeds 227 KB (reduce Q)");
        HpvLaunch l; memset(&l, 0, sizeof(l));
        long long out = 0;
        l.kind = HPV_K_VARFWD; l.op = 1; l.block = HPV_THREADS; l.smem = c->fwd_smem; l.out = &out;
        HPV_CK(hpv_dispatch(k, l));
        if (out < 1) return fail(c, HPV_ERR_LIMIT, "forward kernel cannot be resident on an SM");
        c->fwd_ctas_per_sm = (int)out;
    }
    const int rows = (c->net.dim == 2) ? c->Q : 1;
    // tiles of one MMA tile (128 points) for the tensor-core form: 10 or 11 per CTA at C3 instead of 5 or 6 of 256
    c->part_tile = c->fwd_tc_active ? HPV_TC_MTILE : HPV_FWD_TILE;
    {
        // cost model of the tensor-core form (fitted to per-CTA phase timings at C3, profiles/round2_tuning/
        // varfwd_tc_cta_timeline.txt): an element boundary inside a CTA's range costs about 2.4 tiles, the second CTA of
        // an SM runs its tiles 16 % slower than the first.  HPV_FWD_BALANCE=0 switches the weighting off.
        HpvPartitionCost pc;
        const char* ev = getenv("HPV_FWD_BALANCE");
        if (c->fwd_tc_active && !(ev && atoi(ev) == 0)) {
            const long long ntiles = (long long)c->n_el * ((rows * c->Q + c->part_tile - 1) / c->part_tile);
            const double avg = (double)ntiles / (double)(c->n_sm * c->fwd_ctas_per_sm);
            pc.cross = avg >= 8.0 ? 2.4 : 0.3 * avg;
            if (c->fwd_ctas_per_sm == 2) { pc.first_wave = c->n_sm; pc.wave2 = 1.16; }
            if (const char* e1 = getenv("HPV_FWD_BALANCE_CROSS")) pc.cross = atof(e1);          // tuning overrides
            if (const char* e2 = getenv("HPV_FWD_BALANCE_WAVE2")) { if (c->fwd_ctas_per_sm == 2) pc.wave2 = atof(e2); }
        }
        hpv_partition(c->part, c->n_el, rows * c->Q, c->part_tile, c->n_sm * c->fwd_ctas_per_sm, c->fwd_tc_active ? 0 : HPV_THREADS, &pc);
    }
    { int r;
      if ((r = upload(c, c->cta_tile_begin, c->part.cta_tile_begin))) return r;
      if ((r = upload(c, c->el_first_cta, c->part.el_first_cta))) return r;
      if ((r = upload(c, c->el_part_off, c->part.el_part_off))) return r;
      if ((r = upload(c, c->el_nparts, c->part.el_nparts))) return r; }
    HPV_CK(c->Upart.alloc((size_t)c->part.total_parts * HPV_NP * HPV_NP));
    HPV_CK(c->counters.alloc(c->n_el + 1));
    HPV_CK(cudaMemsetAsync(c->counters.p, 0, (c->n_el + 1) * sizeof(unsigned int), c->stream));
    HPV_CK(c->Res.alloc((size_t)c->n_el * c->nty * c->ntx));
    HPV_CK(c->el_loss.alloc(c->n_el));
    HPV_CK(c->loss.alloc(1));

    // backward launch plan
    const long long npts = (long long)c->n_el * rows * c->Q;
    { int r = plan_bwd(c, bwd_key_of(c), npts, c->bwd_block, c->bwd_grid, c->bwd_smem); if (r) return r; }
    c->bwd_ctas_per_sm = (c->bwd_grid + c->n_sm - 1) / c->n_sm;
    c->bwd_tc_active = false; c->bwd_tcw_active = false;
    {
        // tensor-core form of the reverse sweep (hpv_varbwd_tc.cuh) when its TMEM / shared-memory plan fits
        const HpvKernelKey kb = bwd_key_of(c);
        const int nch_b = kb.dir ? 2 : hpv_mode_nch(kb.dim, kb.mx, kb.my);
        const bool want_tc = c->bwd_tc == 1 || (c->bwd_tc < 0 && npts <= (long long)64 * c->n_sm);
        if (want_tc && hpv_tc_supported(nch_b, kb.hp)) {
            HpvVarArgs va; memset(&va, 0, sizeof(va));
            va.nhid = c->net.nhid;
            HpvBwdArgs bb; memset(&bb, 0, sizeof(bb)); bb.v = va;
            HpvLaunch l; memset(&l, 0, sizeof(l));
            long long out = 0;
            l.kind = HPV_K_MLPBWD_TC; l.bwd = &bb; l.out = &out; l.block = HPV_THREADS; l.wg = c->bwd_tcw ? 1 : 0;
            l.op = 5;
            HPV_CK(hpv_dispatch(kb, l));
            c->bwd_tcw_active = out != 0;
            l.op = 2;
            HPV_CK(hpv_dispatch(kb, l));
            const size_t sm_tc = (size_t)out;
            if (sm_tc <= 227 * 1024) {
                l.op = 1; l.smem = sm_tc;
                HPV_CK(hpv_dispatch(kb, l));
                if (out >= 1) {
                    const long long n_tiles = (npts + HPV_TC_MTILE - 1) / HPV_TC_MTILE;
                    const long long slots = (long long)c->n_sm * out;
                    c->bwd_tc_active = true; c->bwd_smem = sm_tc; c->bwd_block = HPV_THREADS; c->bwd_ctas_per_sm = (int)out;
                    c->bwd_grid = (int)(n_tiles < slots ? (n_tiles < 1 ? 1 : n_tiles) : slots);
                }
            }
        }
    }
    { int r = ensure_grad_buffers(c, c->bwd_grid); if (r) return r; }
    HPV_CK(c->Gbar.alloc((size_t)c->form.n_terms * npts));
    c->slabs_per_el = (rows + HPV_ADJ_RS - 1) / HPV_ADJ_RS;
    c->adj_grid = c->n_el * c->slabs_per_el;
    c->adj_smem = (size_t)hpv_adj_smem(a).total * 4;
    if (c->adj_smem > 227 * 1024) return fail(c, HPV_ERR_LIMIT, "adjoint projection shared-memory plan exceeds 227 KB");
    c->ready = true;
    return HPV_OK;
}

// defer_total: the sum of the element losses is left to the loss assembly of the step's gradient reduction
// (hpv_losses_warp) instead of the forward kernel's very last CTA -- two fences and an atomic off the critical path of
// every element's last CTA
int launch_forward(hpv_ctx* c, bool defer_total = false) {
    // the tensor-core form reads the parameters from global memory (theta_pad), not from the constant-memory mirror
    if (!c->fwd_tc_active) { int r = refresh_mirror(c, HPV_K_VARFWD); if (r) return r; }
    HpvVarArgs a; fill_var_args(c, a);
    a.defer_total = defer_total ? 1 : 0;
    HpvLaunch l; memset(&l, 0, sizeof(l));
    l.kind = c->fwd_tc_active ? HPV_K_VARFWD_TC : HPV_K_VARFWD; l.op = 0; l.grid = c->part.n_ctas; l.block = HPV_THREADS; l.smem = c->fwd_smem;
    l.stream = c->stream; l.var = &a;
    HPV_CK(hpv_dispatch(key_of(c, c->form.mx, c->form.my), l));
    c->launches += 1;
    return HPV_OK;
}

int launch_adjproj(hpv_ctx* c) {
    HpvAdjArgs aa; fill_var_args(c, aa.v);
    aa.Gbar = c->Gbar.p; aa.slabs_per_el = c->slabs_per_el;
    HPV_CK(hpv_launch_adjproj(aa, c->adj_grid, c->adj_smem, c->stream));
    c->launches += 1;
    return HPV_OK;
}

int launch_mlpbwd_var(hpv_ctx* c) {
    // the tensor-core form reads the parameters from global memory (theta_pad), not from the constant-memory mirror
    if (!c->bwd_tc_active) { int r = refresh_mirror(c, HPV_K_MLPBWD); if (r) return r; }
    HpvBwdArgs ba; fill_var_args(c, ba.v);
    const int rows = (c->net.dim == 2) ? c->Q : 1;
    ba.Gbar = c->Gbar.p; ba.n_points = c->n_el * rows * c->Q;
    ba.pts = nullptr;
    ba.stagger_ns = c->bwd_stagger_ns;
    HpvLaunch l; memset(&l, 0, sizeof(l));
    l.kind = c->bwd_tc_active ? HPV_K_MLPBWD_TC : HPV_K_MLPBWD; l.op = 0; l.grid = c->bwd_grid; l.block = c->bwd_block; l.smem = c->bwd_smem;
    l.stream = c->stream; l.bwd = &ba; l.wg = c->bwd_tcw ? 1 : 0;
    HPV_CK(hpv_dispatch(bwd_key_of(c), l));
    c->launches += 1;
    return HPV_OK;
}

// `la` non-null: the launch also assembles the loss values (one extra CTA), see hpv_gradreduce_kernel.
int launch_gradreduce(hpv_ctx* c, int n_parts, int accumulate, const HpvLossArgs* la = nullptr, const HpvAdamArgs* adam = nullptr,
                      bool exchange = false) {
    HpvGradReduceArgs g;
    g.grad_part = c->grad_part.p; g.n_parts = n_parts; g.stride = c->grad_stride; g.n = c->net.theta_pad_n + 1;
    g.grad_pad = c->redbuf.p; g.accumulate = accumulate;
    HpvPeerArgs pa;
    if (exchange) {
        memset(&pa, 0, sizeof(pa));
        pa.nranks = c->peer_n; pa.rank = c->peer_rank; pa.seq = ++c->peer_seq; pa.nvp = c->peer_nvp; pa.nchunks = c->peer_nchunks;
        for (int r = 0; r < c->peer_n; ++r) { pa.inbox[r] = c->peer_inbox[r]; pa.flags[r] = c->peer_flags[r]; }
        pa.err = c->peer_err.p; pa.timeout_ns = c->peer_timeout_ns;
    }
    HPV_CK(hpv_launch_gradreduce(g, la, adam, exchange ? &pa : nullptr, c->loss_off, c->stream));
    c->launches += 1;
    return HPV_OK;
}

int launch_points(hpv_ctx* c, int mx, int my, int n, const float* pts, float* u, float* d1, float* d2,
                  PointSet* ps, bool want_adjoint) {
    { int r = refresh_mirror(c, HPV_K_POINTS); if (r) return r; }
    HpvPointArgs p; memset(&p, 0, sizeof(p));
    p.theta_pad = c->theta_pad.p; p.theta_pad_n = c->net.theta_pad_n; p.nhid = c->net.nhid; p.eps = c->eps.p;
    p.off_wo = hpv_off_wo(c->net.dim, c->net.hp, c->net.nhid);

    p.n = n; p.pts = pts; p.out_u = u; p.out_d1 = d1; p.out_d2 = d2;
    int grid = (n + HPV_THREADS - 1) / HPV_THREADS;
    if (grid > c->n_sm * 4) grid = c->n_sm * 4;
    if (grid < 1) grid = 1;
    float* gbar = nullptr;
    if (ps) {
        for (int f = 0; f < HPV_NFIELDS; ++f) { p.a0[f] = ps->a0[f]; p.a1[f] = ps->a1[f]; }
        p.target = ps->target.p; p.weight = ps->weight; p.resid = ps->resid.p; p.blk_loss = ps->blk_loss.p;
        ps->n_ctas = grid;
        if (want_adjoint) gbar = ps->gbar.p;
    }
    p.n_ctas = grid;
    HpvLaunch l; memset(&l, 0, sizeof(l));
    l.kind = HPV_K_POINTS; l.op = 0; l.grid = grid; l.block = HPV_THREADS;
    {
        const HpvKernelKey kk = key_of(c, mx, my);
        l.smem = (size_t)(HPV_THREADS + hpv_slot_floats(kk.dim, kk.mx, kk.my, kk.hp, HPV_THREADS)) * 4;
    }
    l.stream = c->stream; l.pts = &p; l.gbar_out = gbar;
    HPV_CK(hpv_dispatch(key_of(c, mx, my), l));
    c->launches += 1;
    return HPV_OK;
}

int launch_mlpbwd_points(hpv_ctx* c, PointSet& ps, int& grid_out) {
    int block = 0, grid = 0; size_t smem = 0;
    { int r = plan_bwd(c, key_of(c, ps.mx, ps.my), ps.n, block, grid, smem); if (r) return r; }
    { int r = refresh_mirror(c, HPV_K_MLPBWD); if (r) return r; }
    HpvBwdArgs ba; memset(&ba, 0, sizeof(ba));
    HpvVarArgs& a = ba.v;
    a.theta_pad = c->theta_pad.p; a.theta_pad_n = c->net.theta_pad_n; a.nhid = c->net.nhid; a.eps = c->eps.p;
    a.off_wo = hpv_off_wo(c->net.dim, c->net.hp, c->net.nhid);

    a.Q = 1; a.rows = 1; a.n_terms = 1;
    a.terms[0] = hpv_term_zero();
    for (int f = 0; f < HPV_NFIELDS; ++f) { a.terms[0].a0[f] = ps.a0[f]; a.terms[0].a1[f] = ps.a1[f]; }
    a.grad_part = c->grad_part.p; a.grad_stride = c->grad_stride;
    ba.Gbar = ps.gbar.p; ba.n_points = ps.n; ba.pts = ps.pts.p;
    if ((size_t)grid * c->grad_stride > c->grad_part.n) grid = (int)(c->grad_part.n / c->grad_stride);
    HpvLaunch l; memset(&l, 0, sizeof(l));
    l.kind = HPV_K_MLPBWD; l.op = 0; l.grid = grid; l.block = block; l.smem = smem; l.stream = c->stream; l.bwd = &ba;
    HPV_CK(hpv_dispatch(key_of(c, ps.mx, ps.my), l));
    c->launches += 1;
    grid_out = grid;
    return HPV_OK;
}

// total = wv*lossv + point losses, for the case that no gradient reduction ran (nothing to fuse it into)
__global__ void hpv_losses_kernel(const HpvLossArgs a) { hpv_losses_warp(a, threadIdx.x); }

// After a synchronisation: did a wait of the peer exchange time out (a peer rank died or never launched)?
int check_peer_error(hpv_ctx* c) {
    if (!c->peer_n) return HPV_OK;
    unsigned e = 0;
    HPV_CK(cudaMemcpy(&e, c->peer_err.p, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) return fail(c, HPV_ERR_CUDA, "peer gradient exchange timed out (HPV_PEER_TIMEOUT_S, default 20 s): a rank of the node did not arrive");
    return HPV_OK;
}

int need_net(hpv_ctx* c) {
    if (!c) return HPV_ERR_ARG;
    if (!c->have_net) return fail(c, HPV_ERR_STATE, "hpv_set_network has not been called");
    return HPV_OK;
}

}  // namespace

extern "C" {

int hpv_abi_version(void) { return 1; }

int hpv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int hpv_create(hpv_ctx** out, int device) {
    hpv_ctx* c = nullptr;
    if (!out) return fail(c, HPV_ERR_ARG, "ctx output pointer is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(c, HPV_ERR_CUDA, "no CUDA device: libhpv has no CPU path");
    }
    if (device < 0 || device >= n) return fail(c, HPV_ERR_ARG, "device index out of range");
    if (device >= 16) return fail(c, HPV_ERR_LIMIT, "device index >= 16 (the per-device launch caches hold 16 entries)");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(c, HPV_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(c, HPV_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10) return fail(c, HPV_ERR_CUDA, "libhpv is built for sm_100a (B200) only");
    hpv_ctx* ctx = new hpv_ctx();
    ctx->device = device; ctx->n_sm = prop.multiProcessorCount;
    if (const char* ev = getenv("HPV_BWD_DIR")) ctx->bwd_dir = atoi(ev) != 0;
    if (const char* ev = getenv("HPV_ADAM_DIRECT")) ctx->adam_direct = atoi(ev) != 0;
    if (const char* ev = getenv("HPV_FWD_TC")) ctx->fwd_tc = atoi(ev) != 0;
    if (const char* ev = getenv("HPV_GRAPH")) ctx->use_graph = atoi(ev) != 0;
    if (const char* ev = getenv("HPV_BWD_TC")) ctx->bwd_tc = atoi(ev) != 0 ? 1 : 0;
    if (const char* ev = getenv("HPV_BWD_TCW")) ctx->bwd_tcw = atoi(ev) != 0;
    if (const char* ev = getenv("HPV_BWD_STAGGER_NS")) ctx->bwd_stagger_ns = atoi(ev);
    if (const char* ev = getenv("HPV_PEER_TIMEOUT_S")) { const double v = atof(ev); if (v > 0) ctx->peer_timeout_ns = (unsigned long long)(v * 1e9); }
    e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete ctx;
        return fail(c, HPV_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return HPV_OK;
}

void hpv_destroy(hpv_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->theta_pad.release(); c->eps.release(); c->master.release(); c->adam_m.release(); c->adam_v.release();
    c->grad_out.release(); c->pad_index.release(); c->pad_index2.release(); c->ref_index.release(); c->adam_clock.release(); c->xi1.release();
    for (int t = 0; t < HPV_NTAB; ++t) { c->tab[t].release(); c->tabN[t].release(); }
    c->geom.release(); c->F.release(); c->Res.release(); c->el_loss.release(); c->Upart.release(); c->Gbar.release();
    c->ntest.release(); c->cta_tile_begin.release(); c->el_first_cta.release(); c->el_part_off.release();
    c->el_nparts.release(); c->counters.release(); c->loss.release(); c->grad_part.release(); c->redbuf.release();
    for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) {
        c->ps[s].pts.release(); c->ps[s].target.release(); c->ps[s].resid.release(); c->ps[s].gbar.release();
        c->ps[s].blk_loss.release();
    }
    c->param_blob.release();
    for (int i = 0; i < 2; ++i) if (c->step_graph[i]) cudaGraphExecDestroy(c->step_graph[i]);
    c->hist_ring.release(); c->hist_t0.release();
    if (c->param_stage) cudaFreeHost(c->param_stage);
    if (c->param_staged) cudaEventDestroy(c->param_staged);
    for (void* q : c->peer_opened) cudaIpcCloseMemHandle(q);
    c->peer_inbox_own.release(); c->peer_flags_own.release(); c->peer_err.release();
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    {
        std::lock_guard<std::mutex> lock(g_owner_mutex);
        for (int h = 0; h < 3; ++h) for (int k = 0; k < 3; ++k) if (g_owner[c->device & 15][h][k] == c) g_owner[c->device & 15][h][k] = nullptr;
    }
    delete c;
}

const char* hpv_last_error(hpv_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int hpv_set_stream(hpv_ctx* c, void* s) {
    if (!c) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    HPV_CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    c->epoch++;
    return HPV_OK;
}

int hpv_sync(hpv_ctx* c) {
    if (!c) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    HPV_CK(cudaStreamSynchronize(c->stream));
    return HPV_OK;
}

int hpv_set_network(hpv_ctx* c, int dim, const int* layers, int n_layers, int act) {
    if (!c || !layers) return fail(c, HPV_ERR_ARG, "NULL argument");
    HPV_CK(cudaSetDevice(c->device));
    std::string err;
    HpvNet net;
    if (!hpv_net_setup(net, dim, layers, n_layers, act, err)) return fail(c, HPV_ERR_ARG, err);
    if (net.theta_pad_n > HPV_CTHETA_MAX) return fail(c, HPV_ERR_LIMIT, "network too large for the constant-memory parameter mirror (12288 padded floats incl. transposed copies)");
    if (c->peer_n || c->peer_inbox_own.p)
        return fail(c, HPV_ERR_STATE, "hpv_set_network after hpv_peer_export: the peer exchange buffers are sized for the network they were exported with");
    HPV_CK(cudaStreamSynchronize(c->stream));                        // nothing in flight may still use the old buffers
    c->net = net; c->have_net = true; c->ready = false; c->epoch++; c->grad_ready = false;
    c->mirror[0] = c->mirror[1] = c->mirror[2] = nullptr;
    theta_changed(c);
    const int P = net.n_theta;
    HPV_CK(c->theta_pad.alloc(net.theta_pad_n));
    HPV_CK(cudaMemsetAsync(c->theta_pad.p, 0, net.theta_pad_n * sizeof(float), c->stream));
    HPV_CK(c->eps.alloc(1));
    HPV_CK(cudaMemsetAsync(c->eps.p, 0, sizeof(float), c->stream));
    HPV_CK(c->master.alloc(P + 1)); HPV_CK(c->adam_m.alloc(P + 1)); HPV_CK(c->adam_v.alloc(P + 1));
    HPV_CK(c->grad_out.alloc(P + 1 + 8));
    HPV_CK(cudaMemsetAsync(c->master.p, 0, (P + 1) * sizeof(double), c->stream));
    HPV_CK(c->adam_clock.alloc(6));
    { int r = upload(c, c->pad_index, net.pad_index); if (r) return r; }
    { int r = upload(c, c->pad_index2, net.pad_index2); if (r) return r; }
    {
        std::vector<int> ref(net.theta_pad_n + 1, -1);           // padded index -> reference-order index
        for (int i = 0; i < P; ++i) ref[net.pad_index[i]] = i;
        ref[net.theta_pad_n] = P;                                 // the eps slot of the reduce buffer
        int r = upload(c, c->ref_index, ref); if (r) return r;
    }
    for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) c->ps[s].active = false;
    return hpv_reset_optimizer(c);
}

int hpv_num_params(hpv_ctx* c) {
    if (!c || !c->have_net) return HPV_ERR_STATE;
    return c->net.n_theta;
}

int hpv_set_params(hpv_ctx* c, const double* theta, int n, double eps) {
    { int r = need_net(c); if (r) return r; }
    if (!theta || n != c->net.n_theta) return fail(c, HPV_ERR_ARG, "theta must hold hpv_num_params() values");
    HPV_CK(cudaSetDevice(c->device));
    std::vector<float>& pad = c->host_f32;
    hpv_pad_theta(c->net, theta, pad);
    pad.push_back((float)eps);
    const size_t nb_pad = pad.size() * sizeof(float), off_m = (nb_pad + 15) & ~(size_t)15, nb_m = (size_t)(n + 1) * sizeof(double);
    if (c->param_stage_bytes < off_m + nb_m) {
        if (c->param_stage) { HPV_CK(cudaStreamSynchronize(c->stream)); cudaFreeHost(c->param_stage); c->param_stage = nullptr; }
        HPV_CK(cudaMallocHost(&c->param_stage, off_m + nb_m));
        c->param_stage_bytes = off_m + nb_m;
        if (!c->param_staged) HPV_CK(cudaEventCreateWithFlags(&c->param_staged, cudaEventDisableTiming));
        c->param_stage_busy = false;
    }
    if (c->param_stage_busy) HPV_CK(cudaEventSynchronize(c->param_staged));       // previous upload consumed?
    float* sp = reinterpret_cast<float*>(c->param_stage);
    double* sm = reinterpret_cast<double*>(c->param_stage + off_m);
    memcpy(sp, pad.data(), nb_pad);
    memcpy(sm, theta, (size_t)n * sizeof(double)); sm[n] = eps;
    // one host-to-device copy of the blob, one kernel that scatters it into the parameter buffers and into the
    // constant-memory mirrors this context owns (instead of three copies now and one re-staging copy per mirror later)
    if (c->param_blob.n < off_m + nb_m) HPV_CK(c->param_blob.alloc(off_m + nb_m));
    HPV_CK(cudaMemcpyAsync(c->param_blob.p, c->param_stage, off_m + nb_m, cudaMemcpyHostToDevice, c->stream));
    HPV_CK(cudaEventRecord(c->param_staged, c->stream));
    c->param_stage_busy = true;
    HpvParamScatterArgs pa; memset(&pa, 0, sizeof(pa));
    pa.blob_pad = reinterpret_cast<const float*>(c->param_blob.p);
    pa.blob_master = reinterpret_cast<const double*>(c->param_blob.p + off_m);
    pa.theta_pad_n = c->net.theta_pad_n; pa.n_master = n + 1;
    pa.theta_pad = c->theta_pad.p; pa.eps = c->eps.p; pa.master = c->master.p;
    bool wrote[3] = {false, false, false};
    if (c->adam_direct) {
        std::lock_guard<std::mutex> lock(g_owner_mutex);
        const int hpi = c->net.hp == 8 ? 0 : (c->net.hp == 20 ? 1 : 2);
        for (int k = 0; k < 3; ++k)
            if (c->mirror[k] && g_owner[c->device & 15][hpi][k] == c) { pa.mirror[k] = c->mirror[k]; wrote[k] = true; }
    }
    HPV_CK(hpv_launch_param_scatter(pa, c->stream));
    for (int k = 0; k < 3; ++k) c->mirror_stale[k] = !wrote[k];      // a written mirror now holds the complete new set
    return HPV_OK;
}

int hpv_get_params(hpv_ctx* c, double* theta, int n, double* eps) {
    { int r = need_net(c); if (r) return r; }
    if (!theta || n != c->net.n_theta) return fail(c, HPV_ERR_ARG, "theta must hold hpv_num_params() values");
    HPV_CK(cudaSetDevice(c->device));
    std::vector<double>& m = c->host_f64;
    m.resize(n + 1);
    HPV_CK(cudaMemcpyAsync(m.data(), c->master.p, (n + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));
    memcpy(theta, m.data(), n * sizeof(double));
    if (eps) *eps = m[n];
    return HPV_OK;
}

int hpv_set_quadrature(hpv_ctx* c, int Q, const double* xi, const double* w) {
    if (!c || !xi || !w) return fail(c, HPV_ERR_ARG, "NULL argument");
    if (Q < 2 || Q > HPV_QMAX) return fail(c, HPV_ERR_LIMIT, "Q must be in [2, 128]");
    c->Q = Q; c->xi.assign(xi, xi + Q); c->w.assign(w, w + Q);
    c->have_quad = true; c->have_tabs = false; c->ready = false; c->epoch++;
    return HPV_OK;
}

int hpv_set_test_tables(hpv_ctx* c, int N, const double* T, const double* D1, const double* D2, const double* d1b) {
    if (!c || !T) return fail(c, HPV_ERR_ARG, "NULL argument");
    if (!c->have_quad) return fail(c, HPV_ERR_STATE, "hpv_set_quadrature must come first");
    if (N < 1 || N > HPV_NP) return fail(c, HPV_ERR_LIMIT, "N must be in [1, 64]");
    const size_t n = (size_t)N * c->Q;
    c->N = N; c->T.assign(T, T + n);
    if (D1) c->D1.assign(D1, D1 + n); else c->D1.assign(n, 0.0);
    if (D2) c->D2.assign(D2, D2 + n); else c->D2.assign(n, 0.0);
    c->have_d1b = d1b != nullptr;
    if (d1b) c->d1b.assign(d1b, d1b + 2 * (size_t)N);
    c->have_tabs = true; c->ready = false; c->epoch++;
    return HPV_OK;
}

int hpv_set_form(hpv_ctx* c, int problem, int var_form, double V) {
    if (!c) return HPV_ERR_ARG;
    std::string err;
    HpvForm fm;
    if (!hpv_form_setup(fm, problem, var_form, V, err)) return fail(c, HPV_ERR_ARG, err);
    c->form = fm; c->problem = problem; c->var_form = var_form; c->V = V;
    c->have_form = true; c->ready = false; c->epoch++;
    return HPV_OK;
}

int hpv_set_elements(hpv_ctx* c, int n_el, const double* lo, const double* hi, const int* ntest, int ntx, int nty,
                     const double* F_ext) {
    { int r = need_net(c); if (r) return r; }
    if (!lo || !hi) return fail(c, HPV_ERR_ARG, "NULL element corners");
    if (n_el < 1) return fail(c, HPV_ERR_ARG, "n_el must be >= 1 (an empty element batch has no loss)");
    const int dim = c->net.dim;
    if (dim == 1) nty = 1;
    if (ntx < 1 || ntx > HPV_NP || nty < 1 || nty > HPV_NP) return fail(c, HPV_ERR_LIMIT, "ntx, nty must be in [1, 64]");
    HPV_CK(cudaSetDevice(c->device));
    std::vector<float> geom((size_t)n_el * 4);
    std::vector<int> nt((size_t)n_el * 2);
    for (int e = 0; e < n_el; ++e) {
        const double lx = lo[(size_t)e * dim], hx = hi[(size_t)e * dim];
        if (!(hx > lx)) return fail(c, HPV_ERR_ARG, "element with non-positive width");
        geom[4 * e + 0] = (float)lx; geom[4 * e + 1] = (float)((hx - lx) / 2);
        geom[4 * e + 2] = 0.0f; geom[4 * e + 3] = 1.0f;
        if (dim == 2) {
            const double ly = lo[(size_t)e * dim + 1], hy = hi[(size_t)e * dim + 1];
            if (!(hy > ly)) return fail(c, HPV_ERR_ARG, "element with non-positive height");
            geom[4 * e + 2] = (float)ly; geom[4 * e + 3] = (float)((hy - ly) / 2);
        }
        int ex = ntx, ey = nty;
        if (ntest) { ex = ntest[(size_t)e * dim]; ey = (dim == 2) ? ntest[(size_t)e * dim + 1] : 1; }
        if (ex < 1 || ex > ntx || ey < 1 || ey > nty) return fail(c, HPV_ERR_ARG, "per-element ntest out of [1, ntx] x [1, nty]");
        nt[2 * e + 0] = ex; nt[2 * e + 1] = ey;
    }
    c->n_el = n_el; c->ntx = ntx; c->nty = nty; c->has_F = F_ext != nullptr;
    { int r;
      if ((r = upload(c, c->geom, geom))) return r;
      if ((r = upload(c, c->ntest, nt))) return r; }
    const size_t nf = (size_t)n_el * nty * ntx;
    HPV_CK(c->F.alloc(nf));
    if (F_ext) {
        std::vector<float> f(nf);
        for (size_t i = 0; i < nf; ++i) f[i] = (float)F_ext[i];
        int r = upload(c, c->F, f); if (r) return r;
    }
    c->have_el = true; c->ready = false; c->epoch++;
    return HPV_OK;
}

int hpv_update_rhs_f32(hpv_ctx* c, const float* F) {
    if (!c || !F) return fail(c, HPV_ERR_ARG, "NULL argument");
    if (!c->have_el || !c->has_F) return fail(c, HPV_ERR_STATE, "no element batch with a right-hand side is set");
    HPV_CK(cudaSetDevice(c->device));
    HPV_CK(cudaMemcpyAsync(c->F.p, F, (size_t)c->n_el * c->nty * c->ntx * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    return HPV_OK;
}

int hpv_varloss_forward(hpv_ctx* c, double* lossv, float* residual, double* el_loss) {
    if (!c) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    { int r = launch_forward(c); if (r) return r; }
    double lv = 0.0;
    HPV_CK(cudaMemcpyAsync(&lv, c->loss.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (residual)
        HPV_CK(cudaMemcpyAsync(residual, c->Res.p, (size_t)c->n_el * c->nty * c->ntx * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (el_loss) {
        c->host_f32.resize(c->n_el);
        HPV_CK(cudaMemcpyAsync(c->host_f32.data(), c->el_loss.p, c->n_el * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    HPV_CK(cudaStreamSynchronize(c->stream));
    if (lossv) *lossv = lv;
    if (el_loss) for (int e = 0; e < c->n_el; ++e) el_loss[e] = c->host_f32[e];
    return HPV_OK;
}

int hpv_project_field(hpv_ctx* c, const double* field, int ltab, int rtab, double sc, int px, int py, double* out) {
    if (!c || !field || !out) return fail(c, HPV_ERR_ARG, "NULL argument");
    if (ltab < 0 || ltab >= HPV_NTAB || rtab < 0 || rtab >= HPV_NTAB) return fail(c, HPV_ERR_ARG, "table index out of range");
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    const int rows = (c->net.dim == 2) ? c->Q : 1;
    const size_t npts = (size_t)c->n_el * rows * c->Q, nres = (size_t)c->n_el * c->nty * c->ntx;
    std::vector<float> hf(npts);
    for (size_t i = 0; i < npts; ++i) hf[i] = (float)field[i];
    // scratch outputs: the projection must not disturb Res / el_loss / lossv of the variational loss, which a later
    // hpv_varloss_backward consumes
    ScopedBuf<float> dfield, dres, del;
    ScopedBuf<double> dloss;
    HPV_CK(dfield.alloc(npts)); HPV_CK(dres.alloc(nres)); HPV_CK(del.alloc(c->n_el)); HPV_CK(dloss.alloc(1));
    HPV_CK(cudaMemcpyAsync(dfield.p, hf.data(), npts * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    HpvVarArgs a; fill_var_args(c, a);
    a.n_terms = 1;
    a.terms[0] = hpv_term_zero();
    a.terms[0].ltab = ltab; a.terms[0].rtab = rtab; a.terms[0].s = (float)sc; a.terms[0].px = px; a.terms[0].py = py;
    a.F = nullptr; a.field_in = dfield.p;
    a.Res = dres.p; a.el_loss = del.p; a.loss = dloss.p;
    const HpvKernelKey k = key_of(c, 0, 0);
    const HpvFwdSmem fs = hpv_fwd_smem(a, hpv_slot_floats(k.dim, k.mx, k.my, k.hp, HPV_THREADS));
    HpvLaunch l; memset(&l, 0, sizeof(l));
    l.kind = HPV_K_VARFWD; l.op = 0; l.grid = c->part.n_ctas; l.block = HPV_THREADS; l.smem = (size_t)fs.total * 4;
    l.stream = c->stream; l.var = &a;
    cudaError_t e = hpv_dispatch(k, l);
    c->launches += 1;
    std::vector<float> hr(nres);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hr.data(), dres.p, nres * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e2 = cudaStreamSynchronize(c->stream);               // always: the scratch buffers are freed on return
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) return fail(c, HPV_ERR_CUDA, std::string("hpv_project_field: ") + cudaGetErrorString(e));
    for (size_t i = 0; i < nres; ++i) out[i] = hr[i];
    return HPV_OK;
}

int hpv_forward_async(hpv_ctx* c) {
    if (!c) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    return launch_forward(c);
}

// Arguments of the Adam / un-padding kernels.  With update != 0 the caller must call adam_launched() after the launch.
static void fill_adam_args(hpv_ctx* c, HpvAdamArgs& a, int update, bool wrote[3]) {
    memset(&a, 0, sizeof(a));
    a.grad_pad = c->redbuf.p; a.pad_index = c->pad_index.p; a.pad_index2 = c->pad_index2.p; a.n_theta = c->net.n_theta; a.theta_pad_n = c->net.theta_pad_n;
    a.ref_index = c->ref_index.p;
    a.theta = c->master.p; a.m = c->adam_m.p; a.v = c->adam_v.p; a.theta_pad = c->theta_pad.p; a.eps = c->eps.p;
    a.grad_out = c->grad_out.p; a.train_eps = c->train_eps;
    a.lr = c->lr; a.b1 = c->b1; a.b2 = c->b2; a.eps_hat = c->eps_hat;
    a.st_in = c->adam_clock.p + 3 * c->adam_parity; a.st_out = c->adam_clock.p + 3 * (c->adam_parity ^ 1);
    a.update = update;
    wrote[0] = wrote[1] = wrote[2] = false;
    if (update && c->adam_direct) {
        // mirrors this context owns right now take the update in place; a mirror that was already stale stays stale
        std::lock_guard<std::mutex> lock(g_owner_mutex);
        const int hpi = c->net.hp == 8 ? 0 : (c->net.hp == 20 ? 1 : 2);
        for (int k = 0; k < 3; ++k)
            if (c->mirror[k] && g_owner[c->device & 15][hpi][k] == c) { a.mirror[k] = c->mirror[k]; wrote[k] = true; }
    }
}

static void adam_launched(hpv_ctx* c, const bool wrote[3]) {
    c->adam_parity ^= 1;
    c->adam_t += 1.0;
    for (int k = 0; k < 3; ++k) { c->last_wrote[k] = wrote[k]; if (!wrote[k]) c->mirror_stale[k] = true; }
}

static int unpad_grad(hpv_ctx* c, int update) {
    HpvAdamArgs a;
    bool wrote[3];
    fill_adam_args(c, a, update, wrote);
    HPV_CK(hpv_launch_adam(a, c->stream));
    c->launches += 1;
    if (update) adam_launched(c, wrote);
    return HPV_OK;
}

static int read_grad(hpv_ctx* c, double* grad_theta, int n, double* grad_eps) {
    if (!grad_theta || n != c->net.n_theta) return fail(c, HPV_ERR_ARG, "grad_theta must hold hpv_num_params() values");
    { int r = unpad_grad(c, 0); if (r) return r; }
    std::vector<double>& g = c->host_f64;
    g.resize(n + 1);
    HPV_CK(cudaMemcpyAsync(g.data(), c->grad_out.p, (n + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));
    memcpy(grad_theta, g.data(), n * sizeof(double));
    if (grad_eps) *grad_eps = g[n];
    return HPV_OK;
}

int hpv_varloss_backward(hpv_ctx* c, double* grad_theta, int n, double* grad_eps) {
    if (!c) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    const double wv_saved = c->wv;
    c->wv = 1.0;
    int r = launch_adjproj(c);
    if (!r) r = launch_mlpbwd_var(c);
    if (!r) r = launch_gradreduce(c, c->bwd_grid, 0);
    c->wv = wv_saved;
    if (r) return r;
    return read_grad(c, grad_theta, n, grad_eps);
}

int hpv_net_u(hpv_ctx* c, int n, const double* pts, double* u, double* d1, double* d2) {
    { int r = need_net(c); if (r) return r; }
    if (n < 0 || (n > 0 && !pts)) return fail(c, HPV_ERR_ARG, "bad point array");
    if (n == 0) return HPV_OK;
    HPV_CK(cudaSetDevice(c->device));
    const int dim = c->net.dim;
    std::vector<float> hp((size_t)n * dim);
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = (float)pts[i];
    ScopedBuf<float> dp, du, dd1, dd2;
    HPV_CK(dp.alloc(hp.size()));
    HPV_CK(cudaMemcpyAsync(dp.p, hp.data(), hp.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    if (u) HPV_CK(du.alloc(n));
    if (d1) HPV_CK(dd1.alloc((size_t)n * dim));
    if (d2) HPV_CK(dd2.alloc((size_t)n * dim));
    const int m = d2 ? 2 : (d1 ? 1 : 0);
    int r = launch_points(c, m, m, n, dp.p, du.p, dd1.p, dd2.p, nullptr, false);
    std::vector<float> out;
    auto fetch = [&](DevBuf<float>& b, double* dst, size_t cnt) -> cudaError_t {
        out.resize(cnt);
        cudaError_t e = cudaMemcpyAsync(out.data(), b.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) for (size_t i = 0; i < cnt; ++i) dst[i] = out[i];
        return e;
    };
    cudaError_t e = cudaSuccess;
    if (!r && u) e = fetch(du, u, n);
    if (!r && e == cudaSuccess && d1) e = fetch(dd1, d1, (size_t)n * dim);
    if (!r && e == cudaSuccess && d2) e = fetch(dd2, d2, (size_t)n * dim);
    cudaStreamSynchronize(c->stream);
    if (r) return r;
    if (e != cudaSuccess) return fail(c, HPV_ERR_CUDA, std::string("hpv_net_u: ") + cudaGetErrorString(e));
    return HPV_OK;
}

int hpv_set_point_loss(hpv_ctx* c, int slot, int n, const double* pts, const double* target, const double* a0,
                       const double* a1, double weight) {
    { int r = need_net(c); if (r) return r; }
    if (slot < 0 || slot >= HPV_MAX_POINT_SETS) return fail(c, HPV_ERR_ARG, "slot out of range");
    PointSet& ps = c->ps[slot];
    c->epoch++;
    if (n == 0) { ps.active = false; ps.n = 0; return HPV_OK; }
    if (n < 0 || !pts || !target || !a0) return fail(c, HPV_ERR_ARG, "bad point-loss arguments");
    HPV_CK(cudaSetDevice(c->device));
    const int dim = c->net.dim;
    std::vector<float> hp((size_t)n * dim), ht(n);
    for (size_t i = 0; i < hp.size(); ++i) hp[i] = (float)pts[i];
    for (int i = 0; i < n; ++i) ht[i] = (float)target[i];
    { int r;
      if ((r = upload(c, ps.pts, hp))) return r;
      if ((r = upload(c, ps.target, ht))) return r; }
    HPV_CK(ps.resid.alloc(n)); HPV_CK(ps.gbar.alloc(n)); HPV_CK(ps.blk_loss.alloc((size_t)c->n_sm * 4));
    for (int f = 0; f < HPV_NFIELDS; ++f) { ps.a0[f] = (float)a0[f]; ps.a1[f] = a1 ? (float)a1[f] : 0.0f; }
    if (dim == 1 && (ps.a0[2] != 0 || ps.a0[4] != 0 || ps.a1[2] != 0 || ps.a1[4] != 0))
        return fail(c, HPV_ERR_ARG, "y-derivative fields requested for a 1-D network");
    ps.weight = (float)weight; ps.n = n; ps.mx = 0; ps.my = 0;
    hpv_mode_of_coef(dim, ps.a0, ps.a1, ps.mx, ps.my);
    hpv_canon_mode(dim, ps.mx, ps.my);
    ps.active = true;
    return HPV_OK;
}

int hpv_point_loss_forward(hpv_ctx* c, int slot, double* loss, double* resid) {
    { int r = need_net(c); if (r) return r; }
    if (slot < 0 || slot >= HPV_MAX_POINT_SETS || !c->ps[slot].active) return fail(c, HPV_ERR_ARG, "point-loss slot is not set");
    HPV_CK(cudaSetDevice(c->device));
    PointSet& ps = c->ps[slot];
    { int r = launch_points(c, ps.mx, ps.my, ps.n, ps.pts.p, nullptr, nullptr, nullptr, &ps, false); if (r) return r; }
    std::vector<float>& h = c->host_f32;
    h.resize(ps.n_ctas + ps.n);
    HPV_CK(cudaMemcpyAsync(h.data(), ps.blk_loss.p, ps.n_ctas * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaMemcpyAsync(h.data() + ps.n_ctas, ps.resid.p, ps.n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));
    double acc = 0.0;
    for (int i = 0; i < ps.n_ctas; ++i) acc += h[i];
    if (loss) *loss = acc;
    if (resid) for (int i = 0; i < ps.n; ++i) resid[i] = h[ps.n_ctas + i];
    return HPV_OK;
}

int hpv_configure_training(hpv_ctx* c, double wv, unsigned mask, int train_eps, double lr, double b1, double b2,
                           double eps_hat) {
    if (!c) return HPV_ERR_ARG;
    c->wv = wv; c->mask = mask; c->train_eps = train_eps; c->lr = lr; c->b1 = b1; c->b2 = b2; c->eps_hat = eps_hat;
    c->epoch++;
    return HPV_OK;
}

// fuse_adam: the step's last gradient reduction also applies the Adam update (single-GPU training step).
#define HPV_HIST_CAP 1024
static int loss_and_grad_impl(hpv_ctx* c, bool fuse_adam, bool device_hist = false) {
    { int r = need_net(c); if (r) return r; }
    HPV_CK(cudaSetDevice(c->device));
    const bool use_v = c->wv != 0.0;
    if (use_v) { int r = ensure_ready(c); if (r) return r; }
    else { int r = ensure_grad_buffers(c, 0); if (r) return r; }
    // The gradient reduction of a loss must run before the next reverse sweep overwrites the per-CTA partials;
    // the last one of the step also assembles the loss values (saves a launch).
    int acc = 0, pending = -1;
    if (use_v) {
        int r = launch_forward(c, true);
        if (!r) r = launch_adjproj(c);
        if (!r) r = launch_mlpbwd_var(c);
        if (r) return r;
        pending = c->bwd_grid;
    }
    HpvLossArgs la; memset(&la, 0, sizeof(la));
    la.lossv = c->loss.p; la.wv = (float)c->wv; la.use_v = use_v ? 1 : 0; la.out = c->redbuf.p + c->loss_off;
    if (use_v) { la.el_loss = c->el_loss.p; la.n_el = c->n_el; }     // the forward kernel left the total to the loss assembly
    if (device_hist) {
        la.hist = c->hist_ring.p; la.hist_cap = HPV_HIST_CAP; la.hist_t0 = c->hist_t0.p;
        la.clock = c->adam_clock.p + 3 * c->adam_parity;           // the clock buffer this step's update reads
    }
    for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) {
        if (!(c->mask & (1u << s)) || !c->ps[s].active) continue;
        PointSet& ps = c->ps[s];
        int grid = 0;
        int r = launch_points(c, ps.mx, ps.my, ps.n, ps.pts.p, nullptr, nullptr, nullptr, &ps, true);
        if (!r && pending >= 0) { r = launch_gradreduce(c, pending, acc); acc = 1; pending = -1; }
        if (!r) r = launch_mlpbwd_points(c, ps, grid);
        if (r) return r;
        pending = grid;
        la.blk[s] = ps.blk_loss.p; la.nblk[s] = ps.n_ctas;
    }
    const bool exchange = fuse_adam && c->peer_n > 1;
    if (exchange && pending < 0) {
        // nothing selected on this rank: it still takes part in the exchange, with a zero vector
        HPV_CK(cudaMemsetAsync(c->redbuf.p, 0, c->loss_off * sizeof(float), c->stream));
        pending = 0; acc = 1;
    }
    if (pending >= 0) {
        if (fuse_adam) {
            HpvAdamArgs ad;
            bool wrote[3];
            fill_adam_args(c, ad, 1, wrote);
            int r = launch_gradreduce(c, pending, acc, &la, &ad, exchange);
            if (r) return r;
            adam_launched(c, wrote);
        } else {
            int r = launch_gradreduce(c, pending, acc, &la);
            if (r) return r;
        }
    } else {
        HPV_CK(cudaMemsetAsync(c->redbuf.p, 0, c->loss_off * sizeof(float), c->stream));
        hpv_losses_kernel<<<1, 32, 0, c->stream>>>(la);
        HPV_CK(cudaGetLastError());
        c->launches += 1;
        if (fuse_adam) return unpad_grad(c, 1);
    }
    return HPV_OK;
}

int hpv_loss_and_grad(hpv_ctx* c) { return loss_and_grad_impl(c, false); }

int hpv_reduce_buffer(hpv_ctx* c, void** p, int* n) {
    if (!c || !p || !n) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    *p = c->redbuf.p; *n = c->loss_off + 8;
    return HPV_OK;
}

int hpv_peer_export(hpv_ctx* c, int nranks, unsigned char* handles) {
    if (!c || !handles) return HPV_ERR_ARG;
    if (nranks < 2 || nranks > HPV_MAX_PEERS) return fail(c, HPV_ERR_ARG, "peer exchange supports 2..8 ranks of one node");
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    if (c->peer_n) return fail(c, HPV_ERR_STATE, "peer exchange is already connected");
    c->peer_nchunks = c->loss_off / 32 + 1;
    c->peer_nvp = c->peer_nchunks * 32;
    HPV_CK(c->peer_inbox_own.alloc((size_t)2 * nranks * c->peer_nvp));
    HPV_CK(c->peer_flags_own.alloc((size_t)2 * nranks * c->peer_nchunks));
    HPV_CK(c->peer_err.alloc(1));
    HPV_CK(cudaMemset(c->peer_inbox_own.p, 0, c->peer_inbox_own.n * sizeof(float)));
    HPV_CK(cudaMemset(c->peer_flags_own.p, 0, c->peer_flags_own.n * sizeof(unsigned)));
    HPV_CK(cudaMemset(c->peer_err.p, 0, sizeof(unsigned)));
    cudaIpcMemHandle_t h0, h1;
    HPV_CK(cudaIpcGetMemHandle(&h0, c->peer_inbox_own.p));
    HPV_CK(cudaIpcGetMemHandle(&h1, c->peer_flags_own.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handles, &h0, 64);
    memcpy(handles + 64, &h1, 64);
    return HPV_OK;
}

int hpv_peer_connect(hpv_ctx* c, int rank, int nranks, const unsigned char* all_handles) {
    if (!c || !all_handles) return HPV_ERR_ARG;
    if (nranks < 2 || nranks > HPV_MAX_PEERS || rank < 0 || rank >= nranks) return fail(c, HPV_ERR_ARG, "bad rank / nranks");
    HPV_CK(cudaSetDevice(c->device));
    if (!c->peer_inbox_own.p || (size_t)2 * nranks * c->peer_nvp != c->peer_inbox_own.n)
        return fail(c, HPV_ERR_STATE, "hpv_peer_export(nranks) must be called first");
    if (c->peer_n) return fail(c, HPV_ERR_STATE, "peer exchange is already connected");
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { c->peer_inbox[r] = c->peer_inbox_own.p; c->peer_flags[r] = c->peer_flags_own.p; continue; }
        cudaIpcMemHandle_t h0, h1;
        memcpy(&h0, all_handles + (size_t)r * 128, 64);
        memcpy(&h1, all_handles + (size_t)r * 128 + 64, 64);
        void *q0 = nullptr, *q1 = nullptr;
        HPV_CK(cudaIpcOpenMemHandle(&q0, h0, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened.push_back(q0);
        HPV_CK(cudaIpcOpenMemHandle(&q1, h1, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened.push_back(q1);
        c->peer_inbox[r] = static_cast<float*>(q0); c->peer_flags[r] = static_cast<unsigned*>(q1);
    }
    c->peer_rank = rank; c->peer_n = nranks; c->peer_seq = 0; c->epoch++;
    return HPV_OK;
}

int hpv_adam_step(hpv_ctx* c) {
    { int r = need_net(c); if (r) return r; }
    if (!c->grad_ready) return fail(c, HPV_ERR_STATE, "hpv_loss_and_grad has not been called (for the current network)");
    HPV_CK(cudaSetDevice(c->device));
    return unpad_grad(c, 1);
}

int hpv_read_losses(hpv_ctx* c, double* out, int n) {
    if (!c || !out || n < 1) return HPV_ERR_ARG;
    if (!c->grad_ready) return fail(c, HPV_ERR_STATE, "hpv_loss_and_grad has not been called (for the current network)");
    HPV_CK(cudaSetDevice(c->device));
    float h[8];
    HPV_CK(cudaMemcpyAsync(h, c->redbuf.p + c->loss_off, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));
    { int r = check_peer_error(c); if (r) return r; }
    for (int i = 0; i < n && i < 8; ++i) out[i] = h[i];
    return HPV_OK;
}

int hpv_read_losses_and_grad(hpv_ctx* c, double* losses, int nl, double* g, int n, double* ge) {
    { int r = need_net(c); if (r) return r; }
    if (!losses || nl < 1 || !g || n != c->net.n_theta) return fail(c, HPV_ERR_ARG, "bad output buffers");
    if (!c->grad_ready) return fail(c, HPV_ERR_STATE, "hpv_loss_and_grad has not been called (for the current network)");
    HPV_CK(cudaSetDevice(c->device));
    {   // un-pad the gradient and append the loss values: one kernel, one device-to-host copy, one synchronisation
        HpvAdamArgs a;
        bool wrote[3];
        fill_adam_args(c, a, 0, wrote);
        a.losses_in = c->redbuf.p + c->loss_off;
        HPV_CK(hpv_launch_adam(a, c->stream));
        c->launches += 1;
    }
    std::vector<double>& gh = c->host_f64;
    gh.resize(n + 1 + 8);
    HPV_CK(cudaMemcpyAsync(gh.data(), c->grad_out.p, (n + 1 + 8) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));
    { int r = check_peer_error(c); if (r) return r; }
    memcpy(g, gh.data(), n * sizeof(double));
    if (ge) *ge = gh[n];
    for (int i = 0; i < nl && i < 8; ++i) losses[i] = gh[n + 1 + i];
    return HPV_OK;
}

int hpv_read_grad(hpv_ctx* c, double* g, int n, double* ge) {
    { int r = need_net(c); if (r) return r; }
    if (!c->grad_ready) return fail(c, HPV_ERR_STATE, "hpv_loss_and_grad has not been called (for the current network)");
    HPV_CK(cudaSetDevice(c->device));
    return read_grad(c, g, n, ge);
}

int hpv_reset_optimizer(hpv_ctx* c) {
    { int r = need_net(c); if (r) return r; }
    HPV_CK(cudaSetDevice(c->device));
    const int P = c->net.n_theta;
    HPV_CK(cudaMemsetAsync(c->adam_m.p, 0, (P + 1) * sizeof(double), c->stream));
    HPV_CK(cudaMemsetAsync(c->adam_v.p, 0, (P + 1) * sizeof(double), c->stream));
    const double clock0[6] = {1.0, 1.0, 0.0, 1.0, 1.0, 0.0};
    HPV_CK(cudaMemcpyAsync(c->adam_clock.p, clock0, sizeof(clock0), cudaMemcpyHostToDevice, c->stream));
    HPV_CK(cudaStreamSynchronize(c->stream));                     // clock0 is on the stack
    c->adam_parity = 0;
    c->adam_t = 0.0;
    return HPV_OK;
}

// Capture ONE training step per parity of the optimizer clock (the clock is double-buffered, so the launch arguments
// have period 2) into executable graphs.  The launches are exactly those of the plain path, including the
// programmatic dependent launches between them.  Called after at least two plain steps of the current configuration,
// so that every lazy set-up (launch plans, shared-memory opt-ins, constant-memory mirrors) is behind us and the
// capture contains kernel (and memset/memcpy) nodes only.
static int build_step_graphs(hpv_ctx* c, bool want_hist) {
    for (int i = 0; i < 2; ++i) if (c->step_graph[i]) { cudaGraphExecDestroy(c->step_graph[i]); c->step_graph[i] = nullptr; }
    HPV_CK(cudaStreamSynchronize(c->stream));
    const int parity0 = c->adam_parity;
    const double t_before = c->adam_t;
    const bool stale_before[3] = {c->mirror_stale[0], c->mirror_stale[1], c->mirror_stale[2]};
    int rc = HPV_OK;
    for (int i = 0; i < 2 && rc == HPV_OK; ++i) {
        const int par = parity0 ^ i;
        const long long l0 = c->launches;
        c->adam_parity = par;
        c->used_kinds = 0;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { rc = fail(c, HPV_ERR_CUDA, std::string("cudaStreamBeginCapture: ") + cudaGetErrorString(e)); break; }
        c->capturing = true;
        int r = loss_and_grad_impl(c, true, want_hist);
        c->capturing = false;
        cudaGraph_t g = nullptr;
        e = cudaStreamEndCapture(c->stream, &g);
        c->graph_launches = (int)(c->launches - l0);
        c->launches = l0;
        c->graph_kinds = c->used_kinds;
        for (int k = 0; k < 3; ++k) c->graph_wrote[k] = c->last_wrote[k];
        if (r) { if (g) cudaGraphDestroy(g); rc = r; break; }
        if (e != cudaSuccess) { rc = fail(c, HPV_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e)); break; }
        e = cudaGraphInstantiate(&c->step_graph[par], g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { c->step_graph[par] = nullptr; rc = fail(c, HPV_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    }
    // the captured launches did not run: restore the host-side bookkeeping they advanced
    c->adam_parity = parity0; c->adam_t = t_before;
    for (int k = 0; k < 3; ++k) c->mirror_stale[k] = stale_before[k];
    if (rc != HPV_OK) {
        for (int i = 0; i < 2; ++i) if (c->step_graph[i]) { cudaGraphExecDestroy(c->step_graph[i]); c->step_graph[i] = nullptr; }
        return rc;
    }
    c->graph_epoch = c->epoch; c->graph_hist = want_hist;
    return HPV_OK;
}

static bool graphs_valid(const hpv_ctx* c, bool want_hist) {
    return c->step_graph[0] && c->step_graph[1] && c->graph_epoch == c->epoch && (!want_hist || c->graph_hist);
}

int hpv_train_steps(hpv_ctx* c, int nsteps, double* hist) {
    { int r = need_net(c); if (r) return r; }
    if (nsteps < 0) return fail(c, HPV_ERR_ARG, "nsteps < 0");
    HPV_CK(cudaSetDevice(c->device));
    const bool want_hist = hist && nsteps;
    // Graph replay for single-GPU steps (the peer exchange takes a per-step sequence number from the host).
    const bool graph = c->use_graph && c->peer_n <= 1;
    if (!graph) {
        ScopedBuf<float> dh;
        if (want_hist) HPV_CK(dh.alloc((size_t)nsteps * 6));
        for (int it = 0; it < nsteps; ++it) {
            // one launch sequence per step: forward, adjoint projection, reverse sweep, reduction + losses + Adam
            int r = loss_and_grad_impl(c, true);
            if (!r && want_hist)       // the loss values are those of the parameters the gradient was taken at
                HPV_CK(cudaMemcpyAsync(dh.p + (size_t)it * 6, c->redbuf.p + c->loss_off, 6 * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
            if (r) { cudaStreamSynchronize(c->stream); return r; }
        }
        if (want_hist) {
            std::vector<float> h((size_t)nsteps * 6);
            HPV_CK(cudaMemcpyAsync(h.data(), dh.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            HPV_CK(cudaStreamSynchronize(c->stream));
            for (size_t i = 0; i < h.size(); ++i) hist[i] = h[i];
        }
        return HPV_OK;
    }
    if (want_hist && !c->hist_ring.p) { HPV_CK(c->hist_ring.alloc((size_t)HPV_HIST_CAP * 8)); HPV_CK(c->hist_t0.alloc(1)); }
    if (c->plain_epoch != c->epoch) { c->plain_epoch = c->epoch; c->plain_steps = 0; }
    int done = 0;
    std::vector<float> hh;
    while (done < nsteps) {
        int chunk = nsteps - done;
        if (want_hist && chunk > HPV_HIST_CAP) chunk = HPV_HIST_CAP;
        if (want_hist) {
            const double t0 = c->adam_t;                        // step counter at the first step of this chunk
            HPV_CK(cudaMemcpyAsync(c->hist_t0.p, &t0, sizeof(double), cudaMemcpyHostToDevice, c->stream));
            HPV_CK(cudaStreamSynchronize(c->stream));           // t0 is on the stack
        }
        for (int it = 0; it < chunk; ++it) {
            if (!graphs_valid(c, want_hist) && c->plain_steps >= 2 && chunk - it >= 1) {
                int r = build_step_graphs(c, want_hist || c->graph_hist);
                if (r) return r;
            }
            if (graphs_valid(c, want_hist)) {
                {
                    // the constant-memory copies the step's kernels read must be this context's and current (another
                    // context of the device may have taken them over since the last replay); inside the replays the
                    // optimizer keeps them current
                    for (int k = 0; k < 3; ++k)
                        if ((c->graph_kinds & (1u << k)) || c->mirror[k]) { int r = refresh_mirror(c, k); if (r) return r; }
                }
                HPV_CK(cudaGraphLaunch(c->step_graph[c->adam_parity], c->stream));
                c->launches += c->graph_launches;
                c->adam_parity ^= 1;
                c->adam_t += 1.0;
                for (int k = 0; k < 3; ++k) if (!c->graph_wrote[k]) c->mirror_stale[k] = true;
            } else {
                int r = loss_and_grad_impl(c, true, want_hist);
                if (r) return r;
                c->plain_steps += 1;
            }
        }
        if (want_hist) {
            hh.resize((size_t)chunk * 8);
            HPV_CK(cudaMemcpyAsync(hh.data(), c->hist_ring.p, hh.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            HPV_CK(cudaStreamSynchronize(c->stream));
            for (int i = 0; i < chunk; ++i)
                for (int j = 0; j < 6; ++j) hist[(size_t)(done + i) * 6 + j] = hh[(size_t)i * 8 + j];
        }
        done += chunk;
    }
    return HPV_OK;
}

long long hpv_launch_count(hpv_ctx* c) { return c ? c->launches : 0; }

int hpv_kernel_info(hpv_ctx* c, int* info, int n) {
    if (!c || !info) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    const int v[15] = {c->n_sm, c->part.n_ctas, HPV_THREADS, (int)c->fwd_smem, c->fwd_ctas_per_sm,
                       c->bwd_grid, c->bwd_block, (int)c->bwd_smem, c->bwd_ctas_per_sm,
                       c->adj_grid, (int)c->adj_smem, c->net.hp, bwd_key_of(c).dir, c->fwd_tc_active ? 1 : 0,
                       c->bwd_tc_active ? (c->bwd_tcw_active ? 2 : 1) : 0};
    for (int i = 0; i < n && i < 15; ++i) info[i] = v[i];
    return HPV_OK;
}

int hpv_probe_fp32_peak(hpv_ctx* c, int variant, double* tflops) {
    if (!c || !tflops) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    const int grid = c->n_sm * 8, block = 256, iters = 4096;
    ScopedBuf<float> out;
    HPV_CK(out.alloc((size_t)grid * block));
    cudaEvent_t e0, e1;
    HPV_CK(cudaEventCreate(&e0)); HPV_CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) HPV_CK(hpv_launch_ffma_peak(out.p, grid, block, iters, variant, c->stream));
    const int reps = 10;
    HPV_CK(cudaEventRecord(e0, c->stream));
    for (int i = 0; i < reps; ++i) HPV_CK(hpv_launch_ffma_peak(out.p, grid, block, iters, variant, c->stream));
    HPV_CK(cudaEventRecord(e1, c->stream));
    HPV_CK(cudaEventSynchronize(e1));
    float ms = 0.0f;
    HPV_CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double flops = (double)grid * block * iters * 64.0 * 2.0 * reps;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return HPV_OK;
}

int hpv_time_kernel(hpv_ctx* c, int what, int reps, double* usec) {
    if (!c || !usec || reps < 1) return HPV_ERR_ARG;
    HPV_CK(cudaSetDevice(c->device));
    { int r = ensure_ready(c); if (r) return r; }
    auto run = [&]() -> int {
        if (what == 0) return launch_forward(c, true);        // as the training step launches it
        if (what == 1) return launch_adjproj(c);
        if (what == 2) return launch_mlpbwd_var(c);
        int r = launch_gradreduce(c, c->bwd_grid, 0);
        if (!r) r = unpad_grad(c, 0);
        return r;
    };
    if (what >= 1) { int r = launch_forward(c, true); if (r) return r; }
    if (what >= 2) { int r = launch_adjproj(c); if (r) return r; }
    if (what >= 3) { int r = launch_mlpbwd_var(c); if (r) return r; }
    for (int i = 0; i < 3; ++i) { int r = run(); if (r) return r; }
    cudaEvent_t e0, e1;
    HPV_CK(cudaEventCreate(&e0)); HPV_CK(cudaEventCreate(&e1));
    HPV_CK(cudaEventRecord(e0, c->stream));
    for (int i = 0; i < reps; ++i) { int r = run(); if (r) return r; }
    HPV_CK(cudaEventRecord(e1, c->stream));
    HPV_CK(cudaEventSynchronize(e1));
    float ms = 0.0f;
    HPV_CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *usec = (double)ms * 1e3 / reps;
    return HPV_OK;
}

}  // extern "C"
