/* libhpv -- C ABI of the B200-native hp-VPINN variational-residual engine.
 *
 * The reference (ehsankharazmi/hp-VPINNs) has no FFI: its only interface is the per-script Python class
 * `VPINN` whose constructor builds a TensorFlow-1 graph and whose train() drives `sess.run`.  This header
 * is the boundary a drop-in `VPINN` binds instead of TensorFlow; each entry point names the reference code it
 * replaces (P1D = main/Poisson-1D/hp-VPINN-Poisson-1D.py, P2D = main/Poisson-2D/hp-VPINN-Poisson-2D.py,
 * ADI = main/AdvDiff-Identification/hp-VPINN-AdvDiff-Identification.py).  The ctypes binding is
 * hp-vpinns_b200/_lib.py; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: every function returns 0 on success and a negative code on error (message via
 * hpv_last_error).  Host buffers are borrowed for the duration of the call and are float64 (the reference's
 * dtype at the Python surface) unless stated otherwise; the engine computes in fp32.  A context owns its
 * device memory and is bound to one GPU; it is not thread-safe (one context per GPU / rank).  Calls are
 * ordered on the context's CUDA stream; the ones that return host data synchronise that stream.
 */
#ifndef HPV_H
#define HPV_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct hpv_ctx hpv_ctx;

enum { HPV_OK = 0, HPV_ERR_ARG = -1, HPV_ERR_STATE = -2, HPV_ERR_CUDA = -3, HPV_ERR_LIMIT = -4 };
enum { HPV_SIN = 0, HPV_TANH = 1 };                                   /* P1D:134 (sin); P2D:165, ADI:226 (tanh) */
enum { HPV_PROBLEM_POISSON1D = 0, HPV_PROBLEM_POISSON2D = 1, HPV_PROBLEM_ADVDIFF = 2 };
enum { HPV_FIELD_U = 0, HPV_FIELD_UX = 1, HPV_FIELD_UY = 2, HPV_FIELD_UXX = 3, HPV_FIELD_UYY = 4 };
#define HPV_MAX_POINT_SETS 4

/* Library / device information. */
int hpv_abi_version(void);
int hpv_device_count(void);

/* Context life cycle (replaces tf.Session creation, P2D:133-136, P1D:105-108, ADI:194-197). */
int hpv_create(hpv_ctx** ctx, int device);
void hpv_destroy(hpv_ctx* ctx);
const char* hpv_last_error(hpv_ctx* ctx);       /* ctx may be NULL: error of the last failed hpv_create */
int hpv_set_stream(hpv_ctx* ctx, void* cuda_stream);   /* run on a caller-owned cudaStream_t (NULL: own stream) */
int hpv_sync(hpv_ctx* ctx);

/* Network net_u: layers = [dim, H, ..., H, 1] (initialize_NN, P2D:139-149; neural_net, P2D:158-169). */
int hpv_set_network(hpv_ctx* ctx, int dim, const int* layers, int n_layers, int act);
int hpv_num_params(hpv_ctx* ctx);               /* P: sum of in*out + out over the layers */

/* Parameters in the reference's order: for each layer W (in x out, row-major) then b (out); eps is the
 * AdvDiff diffusivity self.epsilon (ADI:63), ignored by the Poisson problems. */
int hpv_set_params(hpv_ctx* ctx, const double* theta, int n, double eps);
int hpv_get_params(hpv_ctx* ctx, double* theta, int n, double* eps);

/* 1-D Gauss-Lobatto-Jacobi nodes xi in [-1,1] and weights (GaussLobattoJacobiWeights, GJQ:46-61; the 2-D
 * tensor grid of P2D:360-365 is point p = j*Q + i -> (xi[i], xi[j])). */
int hpv_set_quadrature(hpv_ctx* ctx, int Q, const double* xi, const double* w);

/* Test-function tables on the nodes, row-major [N][Q], unweighted: T = Test_fcn (P2D:196-208), D1, D2 =
 * dTest_fcn (P2D:210-229).  d1_bound [N][2] = phi'_n(-1), phi'_n(+1) is needed by Poisson-1D var_form 3 only
 * (P1D:77-80, 89-91), else NULL. */
int hpv_set_test_tables(hpv_ctx* ctx, int N, const double* T, const double* D1, const double* D2,
                        const double* d1_bound);

/* Variational form (the reference reads the globals var_form and V: P1D:82, P2D:93, ADI:161,165). */
int hpv_set_form(hpv_ctx* ctx, int problem, int var_form, double V);

/* Elements: lo/hi [n_el][dim] corners (grid[e], grid[e+1]; P2D:75-79), ntest [n_el][dim] test functions per
 * element and direction or NULL for (ntx, nty) everywhere, F_ext [n_el][nty][ntx] the right-hand side
 * F_ext_total (P2D:384-414, P1D:275-294) or NULL (AdvDiff).  1-D: nty = 1. */
int hpv_set_elements(hpv_ctx* ctx, int n_el, const double* lo, const double* hi, const int* ntest, int ntx,
                     int nty, const double* F_ext);
/* Re-upload the right-hand side of the current element batch from an fp32 host buffer (pinned memory makes
 * the copy asynchronous). */
int hpv_update_rhs_f32(hpv_ctx* ctx, const float* F_ext);

/* Projection of a GIVEN point field onto the test functions with the same fused kernel (no network):
 *   out[e][k][r] = s * Jx^px * Jy^py * sum_p  L[k](eta_p) w_p  field[e][p]  R[r](xi_p) w_p ,
 * ltab/rtab in {0: Test_fcn, 1: first derivative, 2: second derivative, 3: constant 1 (the 1-D "y" direction)}.
 * With field = f_ext at the element's quadrature points, ltab = rtab = 0, s = 1, px = py = 1 this is the
 * right-hand-side assembly F_ext_total of the reference drivers (P2D:384-414; P1D:275-294 with ltab = 3, py = 0).
 * field [n_el][rows*Q] (rows = Q in 2-D, 1 in 1-D, point p = j*Q + i), out [n_el][nty][ntx]. */
int hpv_project_field(hpv_ctx* ctx, const double* field, int ltab, int rtab, double s, int px, int py, double* out);

/* lossv and the element residuals (the loop P2D:68-120 / P1D:64-96 / ADI:108-182 as one fused kernel).
 * residual [n_el][nty][ntx] fp32 and el_loss [n_el] may be NULL. */
int hpv_varloss_forward(hpv_ctx* ctx, double* lossv, float* residual, double* el_loss);
/* The same launch without any host read-back (stream-ordered; results stay on the device). */
int hpv_forward_async(hpv_ctx* ctx);
/* d lossv / d theta (reference order) and d lossv / d eps for the parameters of the last forward
 * (tf.gradients of lossv; what AdamOptimizer.minimize builds, P2D:131-132).  grad_eps may be NULL. */
int hpv_varloss_backward(hpv_ctx* ctx, double* grad_theta, int n, double* grad_eps);

/* net_u and its pure input derivatives at n scattered points [n][dim] (net_u, net_dxu/net_dyu/net_du/net_dtu:
 * P2D:171-185, P1D:140-148, ADI:232-245; predict: P2D:255-257).  u [n], d1 [n][dim], d2 [n][dim]; any may be
 * NULL. */
int hpv_net_u(hpv_ctx* ctx, int n, const double* pts, double* u, double* d1, double* d2);

/* Point-wise least-squares losses:  weight * mean_i (sum_f (a0[f] + eps*a1[f]) field_f(x_i) - target_i)^2.
 * lossb (P2D:122, P1D:98, ADI:184): a0 = {1,0,0,0,0};  lossp / net_f (P2D:123,187-194; P1D:150-155;
 * ADI:185,247-253): the strong-form residual.  slot in [0, HPV_MAX_POINT_SETS). */
int hpv_set_point_loss(hpv_ctx* ctx, int slot, int n, const double* pts, const double* target,
                       const double* a0, const double* a1, double weight);
int hpv_point_loss_forward(hpv_ctx* ctx, int slot, double* loss, double* resid);

/* Total loss = wv * lossv + sum of the active point losses, its gradient, and the optimiser
 * (self.loss / train_op_Adam: P2D:125-132, P1D:100-104, ADI:187-193).
 *   hpv_loss_and_grad : forward + backward of everything selected; leaves [grad | d eps | losses] in the
 *                       device reduce buffer (for the multi-GPU all-reduce) -- no host synchronisation.
 *   hpv_reduce_buffer : device pointer / length (fp32 elements) of that buffer.
 *   hpv_adam_step     : tf.train.AdamOptimizer update (lr_t = lr sqrt(1-b2^t)/(1-b1^t), eps_hat outside the
 *                       bias correction) from the reduce buffer, on the device.
 *   hpv_read_losses   : out[0] total, out[1] lossv, out[2+s] point loss of slot s  (synchronises).
 *   hpv_read_grad     : gradient in reference order from the reduce buffer (synchronises). */
int hpv_configure_training(hpv_ctx* ctx, double wv, unsigned point_slot_mask, int train_eps, double lr,
                           double beta1, double beta2, double eps_hat);
int hpv_loss_and_grad(hpv_ctx* ctx);
int hpv_reduce_buffer(hpv_ctx* ctx, void** dev_ptr, int* n_floats);
int hpv_adam_step(hpv_ctx* ctx);
/* Multi-GPU (one process per GPU of ONE node): the element loss is a sum over elements (varloss_total +=
 * loss_element, P2D:120), so each rank holds a block of elements and the ranks sum [grad | d eps | losses] once per
 * step.  With a peer connection that sum runs INSIDE the step's last kernel over NVLink peer memory (push into
 * every rank's inbox, flags with system-scope release/acquire, sum in rank order => bitwise identical on all
 * ranks), followed by the Adam update in the same launch: hpv_train_steps then needs no NCCL call.
 *   hpv_peer_export  : allocates this rank's inbox + flags, returns their two CUDA IPC handles (128 bytes).
 *   hpv_peer_connect : all_handles = the nranks x 128 bytes gathered from every rank (rank order).
 * Without a connection the caller all-reduces hpv_reduce_buffer itself (NCCL) between hpv_loss_and_grad and
 * hpv_adam_step.  A rank that fails to arrive within 20 s (HPV_PEER_TIMEOUT_S) makes the others report an error instead of hanging. */
int hpv_peer_export(hpv_ctx* ctx, int nranks, unsigned char* handles_128_bytes);
int hpv_peer_connect(hpv_ctx* ctx, int rank, int nranks, const unsigned char* all_handles);
int hpv_read_losses(hpv_ctx* ctx, double* out, int n);
int hpv_read_grad(hpv_ctx* ctx, double* grad_theta, int n, double* grad_eps);
/* Both of the above with ONE host synchronisation (what a host-side optimizer loop needs per step). */
int hpv_read_losses_and_grad(hpv_ctx* ctx, double* losses, int n_losses, double* grad_theta, int n, double* grad_eps);
int hpv_reset_optimizer(hpv_ctx* ctx);
/* nsteps full training steps (loss_and_grad + adam_step) back to back; loss_history [nsteps][6] = total, lossv
 * and the four point losses BEFORE each update (i.e. at the parameters the gradient was taken at); may be NULL. */
int hpv_train_steps(hpv_ctx* ctx, int nsteps, double* loss_history);

/* Measurement helpers for bench.py: launch counter of this context, the dominant kernels' launch geometry
 * (info[0..14] = SMs, forward grid/block/smem/CTAs per SM, reverse-sweep grid/block/smem/CTAs per SM, adjoint
 * projection grid/smem, padded hidden width, 1 if the reverse sweep runs in the directional mode, 1 if the forward
 * kernel runs in its tensor-core form (tcgen05 layer products, TMA-staged tables; HPV_FWD_TC=0 selects the FFMA form), the same for the
 * reverse sweep (HPV_BWD_TC=0)),
 * and the FP32-FFMA probe (variant 0 register operands, 1 constant-bank operand, 2 packed f32x2); the probe
 * returns the achieved TFLOP/s measured with CUDA events on the context's stream. */
long long hpv_launch_count(hpv_ctx* ctx);
int hpv_kernel_info(hpv_ctx* ctx, int* info, int n);
int hpv_probe_fp32_peak(hpv_ctx* ctx, int variant, double* tflops);
/* Time `reps` launches of one kernel group with CUDA events on the context's stream: what = 0 forward,
 * 1 adjoint projection, 2 MLP reverse sweep, 3 gradient reduction + Adam.  Returns mean microseconds. */
int hpv_time_kernel(hpv_ctx* ctx, int what, int reps, double* usec);

#ifdef __cplusplus
}
#endif
#endif
