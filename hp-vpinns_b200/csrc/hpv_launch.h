// Host-side launch interface of the kernels (implemented once per padded hidden width in hpv_kernels_h*.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include "hpv_varbwd.cuh"
#include "hpv_points.cuh"
#include "hpv_varfwd_tc.cuh"
#include "hpv_varbwd_tc.cuh"
#include "hpv_varbwd_tcw.cuh"

// dir != 0 (reverse sweep only): the directional mode HpvMode<2, 1, 0> instead of <2, 1, 1>.
// Launch with programmatic stream serialisation: the kernel's CTAs may start before the previous kernel of the
// stream has finished; the kernel orders itself with hpv_pdl_wait() (hpv_cta.cuh).  HPV_PDL=0 in the environment
// falls back to plain launches.
#if defined(__CUDACC__)
inline bool hpv_pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("HPV_PDL"); on = (e && atoi(e) == 0) ? 0 : 1; }
    return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t hpv_launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid, 1, 1); cfg.blockDim = dim3(block, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = hpv_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

struct HpvKernelKey { int dim, mx, my, hp, act, dir, wg; };   // wg: tensor-core reverse sweep with the weight gradients on the tensor cores too

enum { HPV_K_VARFWD = 0, HPV_K_MLPBWD = 1, HPV_K_POINTS = 2, HPV_K_VARFWD_TC = 3, HPV_K_MLPBWD_TC = 4 };

// op: 0 = launch, 1 = query resident CTAs per SM for (block, smem) into *out, 4 = (reverse sweep) the largest
// block size the kernel was compiled for, 2 = shared memory bytes of the
// MLP reverse sweep for block size `block` into *out, 3 = device address of this translation unit's constant
// parameter slots (hpv_c_theta) into *out.
struct HpvLaunch {
    int kind, op;
    int grid, block;
    size_t smem;
    cudaStream_t stream;
    const HpvVarArgs* var;
    const HpvBwdArgs* bwd;
    const HpvPointArgs* pts;
    float* gbar_out;
    long long* out;
    int wg;                    // reverse sweep, tensor-core form: weight gradients on the tensor cores too (when supported)
};

#define HPV_DECL(hp) \
    cudaError_t hpv_dispatch_h##hp##_fwd(const HpvKernelKey& k, const HpvLaunch& l); \
    cudaError_t hpv_dispatch_h##hp##_bwd(const HpvKernelKey& k, const HpvLaunch& l); \
    cudaError_t hpv_dispatch_h##hp##_pts(const HpvKernelKey& k, const HpvLaunch& l); \
    cudaError_t hpv_dispatch_h##hp##_fwdtc(const HpvKernelKey& k, const HpvLaunch& l); \
    cudaError_t hpv_dispatch_h##hp##_bwdtc(const HpvKernelKey& k, const HpvLaunch& l);
HPV_DECL(8)
HPV_DECL(20)
HPV_DECL(32)
#undef HPV_DECL

inline cudaError_t hpv_dispatch(const HpvKernelKey& k, const HpvLaunch& l) {
#define HPV_CASE(hpv) \
    if (k.hp == hpv) { \
        if (l.kind == HPV_K_VARFWD) return hpv_dispatch_h##hpv##_fwd(k, l); \
        if (l.kind == HPV_K_MLPBWD) return hpv_dispatch_h##hpv##_bwd(k, l); \
        if (l.kind == HPV_K_VARFWD_TC) return hpv_dispatch_h##hpv##_fwdtc(k, l); \
        if (l.kind == HPV_K_MLPBWD_TC) return hpv_dispatch_h##hpv##_bwdtc(k, l); \
        return hpv_dispatch_h##hpv##_pts(k, l); \
    }
    HPV_CASE(8)
    HPV_CASE(20)
    HPV_CASE(32)
#undef HPV_CASE
    return cudaErrorInvalidValue;
}

// mode-independent kernels (hpv_kernels_common.cu)
cudaError_t hpv_launch_adjproj(const HpvAdjArgs& a, int grid, size_t smem, cudaStream_t s);
// Loss assembly: out[0] = wv*lossv + sum of the point losses, out[1] = lossv, out[2+s] = point loss s
// (sum of its per-CTA partials).  One warp.
#define HPV_MAX_POINT_SETS 4
struct HpvLossArgs {
    const double* lossv; float wv; int use_v;
    const float* blk[HPV_MAX_POINT_SETS]; int nblk[HPV_MAX_POINT_SETS];
    float* out;     // [8]: total, lossv, point losses
    // optional device-side loss history (graph-replayed training steps cannot take a per-step destination from the
    // host): the values also go to hist[8 * (t - t0)] with t the optimizer's step counter (clock[2]) -- 0 <= t - t0 < hist_cap
    float* hist; const double* clock; const double* hist_t0; int hist_cap;
    // non-null: lossv is the sum of these element losses (float64, fixed order), formed here instead of in the forward kernel
    const float* el_loss; int n_el;
};
#if defined(__CUDACC__)
__device__ inline void hpv_losses_warp(const HpvLossArgs& a, int lane) {
    float lv = 0.0f;
    if (a.use_v) {
        if (a.el_loss) {
            double acc = 0.0;
            for (int i = lane; i < a.n_el; i += 32) acc += (double)a.el_loss[i];
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            lv = (float)acc;
        } else {
            lv = (float)a.lossv[0];
        }
    }
    float total = a.wv * lv;
    float pl[HPV_MAX_POINT_SETS];
    for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) {
        float acc = 0.0f;
        if (a.blk[s]) for (int i = lane; i < a.nblk[s]; i += 32) acc += a.blk[s][i];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        pl[s] = acc;
        total += acc;
    }
    if (lane == 0) {
        a.out[0] = total; a.out[1] = lv;
        for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) a.out[2 + s] = pl[s];
        if (a.hist) {
            const long long idx = (long long)(a.clock[2] - a.hist_t0[0]);
            if (idx >= 0 && idx < (long long)a.hist_cap) {
                float* h = a.hist + idx * 8;
                h[0] = total; h[1] = lv;
                for (int s = 0; s < HPV_MAX_POINT_SETS; ++s) h[2 + s] = pl[s];
            }
        }
    }
}
#endif
// Gradient exchange between the GPUs of one node over peer memory (NVLink / NVSwitch), fused into the reduction
// kernel.  Every rank owns an inbox  [2 parities][nranks sources][nvp floats]  and arrival flags
// [2][nranks][nchunks] (one per 32-entry chunk of the vector), both mapped into every peer with CUDA IPC.  A CTA
// that has reduced its chunk of the local gradient PUSHES it into every rank's inbox (remote stores), publishes
// the step's sequence number in the flags (release, system scope), waits for the same chunk of every other rank
// in its OWN inbox (local polling, acquire), and sums the nranks copies in rank order -- every rank gets bitwise
// the same sum -- before the Adam update of the chunk's parameters.  Double buffering by the parity of the
// sequence number is enough: a rank can only be one step ahead of a rank it exchanges with.
#define HPV_MAX_PEERS 8
struct HpvPeerArgs {
    int nranks, rank;
    unsigned seq;              // >= 1, the same on every rank, +1 per exchange
    int nvp, nchunks;          // floats per vector (padded to a multiple of 32), chunks of 32 entries
    float* inbox[HPV_MAX_PEERS];
    unsigned* flags[HPV_MAX_PEERS];
    unsigned* err;             // device word, set to 1 when a wait timed out (a peer died): results are invalid
    unsigned long long timeout_ns;
};

// la non-null: one extra CTA assembles the loss values in the same launch.  adam non-null (single-GPU training
// step): every reduced gradient entry is consumed by the Adam update of its parameter in the same launch.
// peer non-null (multi-GPU training step): the reduced vector is exchanged and summed over the ranks first.
cudaError_t hpv_launch_gradreduce(const HpvGradReduceArgs& a, const HpvLossArgs* la, const struct HpvAdamArgs* adam,
                                  const HpvPeerArgs* peer, int loss_off, cudaStream_t s);
struct HpvAdamArgs {
    const float* grad_pad;     // padded gradient (+ d eps at index theta_pad_n)
    const int* pad_index;      // [n_theta] reference-order index -> padded index
    const int* pad_index2;     // [n_theta] second padded location (transposed copy) or -1
    int n_theta, theta_pad_n;
    double* theta;             // [n_theta + 1] float64 master parameters, reference order, eps last
    double* m; double* v;      // Adam moments, same layout
    float* theta_pad;          // padded copy read by the kernels
    float* mirror[3];          // constant-memory mirrors of theta_pad this context currently owns (or null): the
                               // update is written into them directly, saving the device-to-device re-staging copies
    float* eps;                // device scalar read by the kernels
    double* grad_out;          // [n_theta + 1 + 8] unpadded gradient (or null), followed by the loss values
    const float* losses_in;    // [8] loss values to append to grad_out (or null): one device-to-host copy returns both
    int train_eps;
    double lr, b1, b2, eps_hat;
    const int* ref_index;      // [theta_pad_n + 1] padded index -> reference-order index (eps: n_theta), -1: none
    // optimizer clock, double-buffered by the host (st_in = the other buffer of st_out): {beta1^t, beta2^t, t}.
    // The powers are running products, as TF1 keeps them (beta1_power, beta2_power variables).
    const double* st_in;
    double* st_out;
    int update;                // 0: only unpad the gradient, 1: also apply the Adam update
};
cudaError_t hpv_launch_adam(const HpvAdamArgs& a, cudaStream_t s);
// hpv_set_params on the device: one staged blob [theta_pad | eps | (8-byte aligned) master] is scattered into the
// parameter buffers and the constant-memory mirrors the context owns (no separate copies per destination).
struct HpvParamScatterArgs {
    const float* blob_pad;     // [theta_pad_n] padded fp32 parameters, followed by eps
    const double* blob_master; // [n_master] float64 master copy (reference order, eps last)
    int theta_pad_n, n_master;
    float* theta_pad; float* eps; double* master;
    float* mirror[3];
};
cudaError_t hpv_launch_param_scatter(const HpvParamScatterArgs& a, cudaStream_t s);
cudaError_t hpv_launch_ffma_peak(float* out, int grid, int block, int iters, int variant, cudaStream_t s);
