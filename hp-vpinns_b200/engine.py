"""`Engine`: numpy-facing wrapper of one libhpv context (one GPU).  This is the layer the `VPINN` classes in
vpinn.py sit on; it plays the role the TensorFlow session plays in the reference (`self.sess.run(...)`,
P2D:242-243) for the variational-loss path."""
import ctypes

import numpy as np

from . import _lib as L

SIN, TANH = 0, 1
POISSON1D, POISSON2D, ADVDIFF = 0, 1, 2
ACT = {"sin": SIN, "tanh": TANH, SIN: SIN, TANH: TANH}
PROBLEM = {"poisson1d": POISSON1D, "poisson2d": POISSON2D, "advdiff": ADVDIFF,
           POISSON1D: POISSON1D, POISSON2D: POISSON2D, ADVDIFF: ADVDIFF}
U, UX, UY, UXX, UYY = range(5)


class Engine:
    def __init__(self, device=0):
        self._lib = L.load()
        h = ctypes.c_void_p()
        rc = self._lib.hpv_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise L.HpvError("hpv_create failed (%d): %s" % (rc, self._lib.hpv_last_error(None).decode()))
        self._h = h
        self.device = int(device)
        self.dim = None
        self.n_params = 0
        self.n_el = self.ntx = self.nty = 0

    # -- plumbing ---------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise L.HpvError("libhpv error %d: %s" % (rc, self._lib.hpv_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hpv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._lib.hpv_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._ck(self._lib.hpv_sync(self._h))

    # -- set-up -----------------------------------------------------------------------------------------
    def set_network(self, layers, act):
        layers = L.i32(layers)
        self.dim = int(layers[0])
        self.layers = [int(v) for v in layers]
        self._ck(self._lib.hpv_set_network(self._h, self.dim, L.iptr(layers), len(layers), ACT[act]))
        self.n_params = self._lib.hpv_num_params(self._h)

    def set_params(self, theta, eps=0.0):
        theta = L.f64(theta).ravel()
        self._ck(self._lib.hpv_set_params(self._h, L.dptr(theta), theta.size, float(eps)))

    def get_params(self):
        theta = np.zeros(self.n_params)
        eps = ctypes.c_double(0)
        self._ck(self._lib.hpv_get_params(self._h, L.dptr(theta), theta.size, ctypes.byref(eps)))
        return theta, eps.value

    def set_quadrature(self, xi, w):
        xi, w = L.f64(xi).ravel(), L.f64(w).ravel()
        self.Q = xi.size
        self._ck(self._lib.hpv_set_quadrature(self._h, xi.size, L.dptr(xi), L.dptr(w)))

    def set_test_tables(self, T, D1=None, D2=None, d1_bound=None):
        T = L.f64(T)
        N = T.shape[0]
        T = T.reshape(N, -1)
        D1 = L.f64(D1, (N, -1)) if D1 is not None else None
        D2 = L.f64(D2, (N, -1)) if D2 is not None else None
        d1b = L.f64(d1_bound, (N, 2)) if d1_bound is not None else None
        self._ck(self._lib.hpv_set_test_tables(self._h, N, L.dptr(T), L.dptr(D1), L.dptr(D2), L.dptr(d1b)))

    def set_form(self, problem, var_form, V=1.0):
        self._ck(self._lib.hpv_set_form(self._h, PROBLEM[problem], int(var_form), float(V)))

    def set_elements(self, lo, hi, ntx, nty=1, F_ext=None, ntest=None):
        lo = L.f64(lo).reshape(-1, self.dim)
        hi = L.f64(hi).reshape(-1, self.dim)
        n_el = lo.shape[0]
        if self.dim == 1:
            nty = 1
        F = L.f64(F_ext, (n_el, nty, ntx)) if F_ext is not None else None
        nt = L.i32(ntest).reshape(n_el, self.dim) if ntest is not None else None
        self._ck(self._lib.hpv_set_elements(self._h, n_el, L.dptr(lo), L.dptr(hi), L.iptr(nt), int(ntx), int(nty), L.dptr(F)))
        self.n_el, self.ntx, self.nty = n_el, int(ntx), int(nty)

    def update_rhs_f32(self, host_ptr):
        """host_ptr: address of a (pinned) fp32 host buffer [n_el][nty][ntx]."""
        self._ck(self._lib.hpv_update_rhs_f32(self._h, ctypes.c_void_p(host_ptr)))

    # -- right-hand-side assembly (set-up side) -------------------------------------------------
    def project_field(self, field, ltab=0, rtab=0, scale=1.0, px=1, py=1):
        """Projection of a field given at the quadrature points of every element (see hpv_project_field)."""
        rows = self.Q if self.dim == 2 else 1
        field = L.f64(field).reshape(self.n_el, rows * self.Q)
        out = np.zeros((self.n_el, self.nty, self.ntx))
        self._ck(self._lib.hpv_project_field(self._h, L.dptr(field), int(ltab), int(rtab), float(scale), int(px), int(py), L.dptr(out)))
        return out

    def assemble_rhs(self, f_ext, lo, hi, xi):
        """F_ext_total of the reference drivers on the GPU: F[e][k][r] = J sum wx phi_r wy phi_k f(x_p, y_p)
        (P2D:384-414) / F[e][i] = J sum w f(x_p) phi_i (P1D:275-294).  f_ext is the driver's vectorised function."""
        lo = L.f64(lo).reshape(-1, self.dim); hi = L.f64(hi).reshape(-1, self.dim); xi = L.f64(xi).ravel()
        if self.dim == 2:
            x = lo[:, 0, None] + (hi[:, 0, None] - lo[:, 0, None]) / 2 * (xi[None, :] + 1)         # [n_el][Q]
            y = lo[:, 1, None] + (hi[:, 1, None] - lo[:, 1, None]) / 2 * (xi[None, :] + 1)
            field = f_ext(x[:, None, :], y[:, :, None])                                             # [n_el][j][i]
            return self.project_field(field, 0, 0, 1.0, 1, 1)
        x = lo[:, 0, None] + (hi[:, 0, None] - lo[:, 0, None]) / 2 * (xi[None, :] + 1)
        return self.project_field(f_ext(x), 3, 0, 1.0, 1, 0)

    # -- the hot path -------------------------------------------------------------------------------------
    def varloss_forward(self, want_residual=True, want_el_loss=False):
        loss = ctypes.c_double(0)
        res = np.zeros((self.n_el, self.nty, self.ntx), dtype=np.float32) if want_residual else None
        el = np.zeros(self.n_el) if want_el_loss else None
        self._ck(self._lib.hpv_varloss_forward(self._h, ctypes.byref(loss),
                                               res.ctypes.data_as(ctypes.c_void_p) if res is not None else None, L.dptr(el)))
        out = [loss.value]
        if want_residual:
            out.append(res)
        if want_el_loss:
            out.append(el)
        return out[0] if len(out) == 1 else tuple(out)

    def forward_async(self):
        self._ck(self._lib.hpv_forward_async(self._h))

    def varloss_backward(self):
        g = np.zeros(self.n_params)
        ge = ctypes.c_double(0)
        self._ck(self._lib.hpv_varloss_backward(self._h, L.dptr(g), g.size, ctypes.byref(ge)))
        return g, ge.value

    def net_u(self, pts, d1=False, d2=False):
        pts = L.f64(pts).reshape(-1, self.dim)
        n = pts.shape[0]
        u = np.zeros(n)
        a1 = np.zeros((n, self.dim)) if (d1 or d2) else None
        a2 = np.zeros((n, self.dim)) if d2 else None
        self._ck(self._lib.hpv_net_u(self._h, n, L.dptr(pts), L.dptr(u), L.dptr(a1), L.dptr(a2)))
        if d2:
            return u, a1, a2
        if d1:
            return u, a1
        return u

    def set_point_loss(self, slot, pts, target, a0, a1=None, weight=1.0):
        pts = L.f64(pts).reshape(-1, self.dim)
        target = L.f64(target).ravel()
        a0 = L.f64(a0).ravel()
        a1 = L.f64(a1).ravel() if a1 is not None else None
        self._ck(self._lib.hpv_set_point_loss(self._h, int(slot), pts.shape[0], L.dptr(pts), L.dptr(target), L.dptr(a0),
                                              L.dptr(a1), float(weight)))

    def point_loss_forward(self, slot, n):
        loss = ctypes.c_double(0)
        r = np.zeros(n)
        self._ck(self._lib.hpv_point_loss_forward(self._h, int(slot), ctypes.byref(loss), L.dptr(r)))
        return loss.value, r

    # -- training ---------------------------------------------------------------------------------------------
    def configure_training(self, wv=1.0, point_slots=(), train_eps=False, lr=1e-3, beta1=0.9, beta2=0.999, eps_hat=1e-8):
        mask = 0
        for s in point_slots:
            mask |= 1 << int(s)
        self._ck(self._lib.hpv_configure_training(self._h, float(wv), mask, int(bool(train_eps)), float(lr), float(beta1),
                                                  float(beta2), float(eps_hat)))

    def loss_and_grad(self):
        self._ck(self._lib.hpv_loss_and_grad(self._h))

    def reduce_buffer(self):
        p = ctypes.c_void_p()
        n = ctypes.c_int(0)
        self._ck(self._lib.hpv_reduce_buffer(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def peer_export(self, nranks):
        """This rank's inbox/flag IPC handles (bytes) for the peer-memory gradient exchange."""
        buf = ctypes.create_string_buffer(128)
        self._ck(self._lib.hpv_peer_export(self._h, int(nranks), buf))
        return buf.raw

    def peer_connect(self, rank, nranks, all_handles):
        """all_handles: the concatenated peer_export() bytes of every rank, in rank order."""
        if len(all_handles) != 128 * nranks:
            raise ValueError("all_handles must hold 128 bytes per rank")
        self._ck(self._lib.hpv_peer_connect(self._h, int(rank), int(nranks), bytes(all_handles)))

    def adam_step(self):
        self._ck(self._lib.hpv_adam_step(self._h))

    def read_losses(self):
        out = np.zeros(6)
        self._ck(self._lib.hpv_read_losses(self._h, L.dptr(out), out.size))
        return out

    def read_grad(self):
        g = np.zeros(self.n_params)
        ge = ctypes.c_double(0)
        self._ck(self._lib.hpv_read_grad(self._h, L.dptr(g), g.size, ctypes.byref(ge)))
        return g, ge.value

    def read_losses_and_grad(self):
        """(losses[6], grad, d eps) with one host synchronisation."""
        out = np.zeros(6)
        g = np.zeros(self.n_params)
        ge = ctypes.c_double(0)
        self._ck(self._lib.hpv_read_losses_and_grad(self._h, L.dptr(out), out.size, L.dptr(g), g.size, ctypes.byref(ge)))
        return out, g, ge.value

    def reset_optimizer(self):
        self._ck(self._lib.hpv_reset_optimizer(self._h))

    def train_steps(self, nsteps, want_history=True):
        h = np.zeros((nsteps, 6)) if want_history else None
        self._ck(self._lib.hpv_train_steps(self._h, int(nsteps), L.dptr(h)))
        return h

    # -- measurement ------------------------------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.hpv_launch_count(self._h))

    def kernel_info(self):
        v = np.zeros(15, dtype=np.int32)
        self._ck(self._lib.hpv_kernel_info(self._h, L.iptr(v), v.size))
        keys = ["n_sm", "fwd_grid", "fwd_block", "fwd_smem", "fwd_ctas_per_sm", "bwd_grid", "bwd_block", "bwd_smem",
                "bwd_ctas_per_sm", "adj_grid", "adj_smem", "hidden_pad", "bwd_directional", "fwd_tensor_core", "bwd_tensor_core"]
        return dict(zip(keys, (int(x) for x in v)))

    def probe_fp32_peak(self, variant=0):
        t = ctypes.c_double(0)
        self._ck(self._lib.hpv_probe_fp32_peak(self._h, int(variant), ctypes.byref(t)))
        return t.value

    def time_kernel(self, what, reps=20):
        t = ctypes.c_double(0)
        self._ck(self._lib.hpv_time_kernel(self._h, int(what), int(reps), ctypes.byref(t)))
        return t.value
