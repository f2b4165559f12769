// Shared plain-data types between the host side (hpv_api.cu, hpv_host_prep.h) and the kernels.
// Vocabulary follows the reference scripts: elements, quadrature points, test functions, residuals.
#pragma once
#include <stdint.h>

#define HPV_NFIELDS 5          // u, u_x, u_y (u_t), u_xx, u_yy
#define HPV_MAX_TERMS 2        // every var_form of P1D:83-91, P2D:94-115, ADI:162-174 is <= 2 projected terms
#define HPV_NP 64              // padded number of test functions per direction (N <= 64)
#define HPV_THREADS 256        // threads per CTA of the variational kernels
#define HPV_FWD_TILE 256       // granularity of the forward kernel's work partition (points).  32 (warp tiles: every CTA
                               // the same number of points) measured no faster at C3 and 2 % slower at C4
                               // (profiles/r02m): a warp runs the same number of passes either way
#define HPV_CT 8               // max point tiles (of HPV_THREADS points) per chunk: the projection phases (table staging,
                               // two contractions, four barriers) are paid once per chunk
#define HPV_MAX_HIDDEN 8       // hidden layers
#define HPV_QMAX 128           // quadrature points per direction
#define HPV_NTAB 4             // T*w, D1*w, D2*w, ONE
#define HPV_CTHETA_MAX 12288    // floats of padded parameters (incl. transposed copies) in constant memory (48 KB)

// One projected term:  U += s * Jx^px * Jy^py * (L-table) . G . (R-table)^T,
// with the point field  G = sum_f (a0[f] + eps*a1[f]) * field_f  (field order above).
struct HpvTerm {
    float a0[HPV_NFIELDS];
    float a1[HPV_NFIELDS];
    float s;
    int px, py;
    int ltab, rtab;            // 0: T*w, 1: D1*w, 2: D2*w (or the folded boundary table of P1D var_form 3), 3: ONE
};

// Arguments of the fused variational kernels (forward: residual + loss; backward: d loss / d theta, d eps).
struct HpvVarArgs {
    // network (padded layout, see hpv_host_prep.h)
    const float* theta_pad;    // global copy (host emulation; on the device the kernels read the constant-memory mirror)
    int theta_pad_n;
    int nhid;                  // number of hidden layers
    int off_wo;                // offset of the output layer in the padded layout (hpv_off_wo)
    const float* eps;          // device scalar (AdvDiff diffusivity), never null
    // quadrature and test-function tables
    int Q;                     // nodes per direction
    int rows;                  // rows per element: Q in 2-D, 1 in 1-D
    const float* xi1;          // [Q] xi + 1
    const float* tab[HPV_NTAB];// transposed weighted tables [Q][HPV_NP]
    const float* tabN[HPV_NTAB];// the same tables in natural layout [HPV_NP][QP] (+4 floats of padding), QP = align4(Q)
    int QP;
    // elements
    int n_el;
    const float* el_geom;      // [n_el][4] lo_x, halfwidth_x, lo_y, halfwidth_y
    const int* el_ntest;       // [n_el][2] ntx, nty of this element
    int ntx, nty;              // layout sizes of F / Res (max over elements)
    const float* F;            // [n_el][nty][ntx] or null (AdvDiff: Res = U)
    const float* field_in;     // [n_terms][n_el][rows*Q] or null: point fields given instead of the MLP (RHS assembly)
    // terms
    int n_terms;
    HpvTerm terms[HPV_MAX_TERMS];
    // work partition (point tiles of tile_pts points, never straddling elements; tile_pts = HPV_FWD_TILE for the FFMA
    // forward kernel, 128 = one MMA tile for the tensor-core form: finer tiles, better balance over the CTAs)
    int tile_pts;
    int tiles_per_el;
    int n_ctas;
    const int* cta_tile_begin; // [n_ctas+1]
    const int* el_first_cta;   // [n_el]
    const int* el_part_off;    // [n_el]
    const int* el_nparts;      // [n_el]
    // scratch and outputs
    float* Upart;              // [total_parts][HPV_NP][HPV_NP]
    unsigned int* el_done;     // [n_el] arrival counters (self-resetting)
    unsigned int* n_done;      // [1]
    float* Res;                // [n_el][nty][ntx]
    float* el_loss;            // [n_el]
    double* loss;              // [1]
    // backward only
    float* grad_part;          // [n_ctas][grad_stride]
    int grad_stride;           // >= theta_pad_n + 4 (last slot: d eps)
    float* grad_pad;           // [theta_pad_n + 4] reduced gradient in the padded layout
    unsigned int* bwd_done;    // [1]
    float loss_scale;          // multiplies the adjoint (1 for lossv)
    int defer_total;           // != 0: the forward kernel writes el_loss only; loss[0] is not formed (see hpv_losses_warp)
};

// Scattered-point evaluation (net_u and derivatives; boundary / PINN losses).
struct HpvPointArgs {
    const float* theta_pad;
    int theta_pad_n;
    int nhid;
    int off_wo;
    const float* eps;
    int n;
    const float* pts;          // [n][dim]
    float* out_u;              // [n] or null
    float* out_d1;             // [n][dim] or null
    float* out_d2;             // [n][dim] or null
    // point loss:  loss = weight * mean_i (sum_f c[f]*field_f(i) - target_i)^2,  c = a0 + eps*a1
    float a0[HPV_NFIELDS];
    float a1[HPV_NFIELDS];
    const float* target;       // [n]
    float weight;
    float* resid;              // [n] out (field combination minus target) or null
    float* blk_loss;           // [n_ctas]
    float* grad_part;          // [n_ctas][grad_stride] or null (forward only)
    int grad_stride;
    int n_ctas;
};
