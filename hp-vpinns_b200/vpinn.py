"""Drop-in `VPINN` classes: the reference's three per-script classes (P1D:30-224, P2D:27-257, ADI:58-341) with
the SAME constructor signatures, method names and logging, on top of the B200 engine instead of a TensorFlow-1
graph.  `poisson1d.VPINN`, `poisson2d.VPINN`, `advdiff.VPINN` are the names a reference script imports in place
of its own class definition.

The reference classes read script-level globals (var_form, scheme, LR, lossb_weight, V) and append to script
lists (loss_his, total_record).  Here they are keyword arguments; when one is not given, the value is looked up
in the CALLER's module globals (so an unmodified driver body keeps working), then falls back to the reference
script's default.
"""
import sys
import time

import numpy as np

from . import GaussJacobiQuadRule_V3 as GJ
from .engine import Engine

_MISSING = object()


def _from_caller(name, given, default, depth=2):
    if given is not None:
        return given
    g = sys._getframe(depth).f_globals
    v = g.get(name, _MISSING)
    return default if v is _MISSING else v


class _Fetch:
    """Stand-in for a TF tensor handle: `model.sess.run(model.loss)` style read-backs keep working."""

    def __init__(self, name):
        self.name = name


class _Session:
    def __init__(self, model):
        self._m = model

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, (list, tuple)):
            return [self.run(f) for f in fetches]
        if not isinstance(fetches, _Fetch):
            raise TypeError("this engine can only fetch loss/lossb/lossv/lossp/epsilon/train_op_Adam/u_test handles")
        return self._m._fetch(fetches.name)

    def close(self):
        self._m.engine.close()


def xavier_init(size, rng):
    """Xavier truncated-normal draw of xavier_init (P1D:122-126): std = sqrt(2/(in+out)), re-drawn beyond 2 sigma.
    (TensorFlow's random stream is not reproducible outside TensorFlow; the draw comes from numpy.)"""
    std = np.sqrt(2.0 / (size[0] + size[1]))
    w = rng.standard_normal(size)
    bad = np.abs(w) > 2
    while bad.any():
        w[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(w) > 2
    return std * w


class _VPINNBase:
    ACT = "tanh"
    PROBLEM = None

    # ---- network parameters -----------------------------------------------------------------------------
    def initialize_NN(self, layers):
        weights, biases = [], []
        for l in range(len(layers) - 1):
            weights.append(xavier_init([layers[l], layers[l + 1]], self._rng))
            biases.append(np.zeros((1, layers[l + 1])))
        return weights, biases, 0.01           # `a` is a dead variable in the reference (P1D:117)

    def _pack(self):
        return np.concatenate([np.concatenate([W.ravel(), b.ravel()]) for W, b in zip(self.weights, self.biases)])

    def _unpack(self, theta):
        o = 0
        for l in range(len(self.layers) - 1):
            n = self.layers[l] * self.layers[l + 1]
            self.weights[l] = theta[o:o + n].reshape(self.layers[l], self.layers[l + 1]).copy(); o += n
            self.biases[l] = theta[o:o + self.layers[l + 1]].reshape(1, -1).copy(); o += self.layers[l + 1]

    def set_weights(self, weights, biases, epsilon=None):
        self.weights = [np.asarray(W, dtype=np.float64) for W in weights]
        self.biases = [np.asarray(b, dtype=np.float64).reshape(1, -1) for b in biases]
        if epsilon is not None:
            self.epsilon_value = float(np.asarray(epsilon).ravel()[0])
        self.engine.set_params(self._pack(), self.epsilon_value)

    def _pull(self):
        theta, eps = self.engine.get_params()
        self._unpack(theta)
        self.epsilon_value = eps

    # ---- reference method surface ---------------------------------------------------------------------------
    def neural_net(self, X, weights=None, biases=None, a=None):
        """net_u on the GPU for an (n, dim) array (P2D:158-169).  weights/biases are accepted for signature
        compatibility; the engine's current parameters are used."""
        return self.engine.net_u(np.asarray(X, dtype=np.float64)).reshape(-1, 1)

    def Test_fcn(self, N_test, x):
        return GJ.Test_fcn(N_test, x)

    def dTest_fcn(self, N_test, x):
        return GJ.dTest_fcn(N_test, x)

    def _setup_engine(self, device, layers, xi, w, N, lo, hi, ntx, nty, F, ntest, var_form, V=1.0):
        self.engine = Engine(device)
        e = self.engine
        e.set_network(layers, self.ACT)
        e.set_quadrature(xi, w)
        T = GJ.Test_fcn(N, xi)
        D1, D2 = GJ.dTest_fcn(N, xi)
        d1b, _ = GJ.dTest_fcn(N, np.array([-1.0, 1.0]))
        e.set_test_tables(T, D1, D2, d1b)
        e.set_form(self.PROBLEM, var_form, V)
        e.set_elements(lo, hi, ntx, nty, F, ntest)
        e.set_params(self._pack(), self.epsilon_value)
        self.sess = _Session(self)
        self.loss, self.lossb, self.lossv, self.lossp = _Fetch("loss"), _Fetch("lossb"), _Fetch("lossv"), _Fetch("lossp")
        self.train_op_Adam = _Fetch("train_op_Adam")
        self.epsilon = _Fetch("epsilon")

    # ---- multi-GPU (one process per GPU, element shards) -------------------------------------------
    def _init_sharding(self, elements, owns_point_losses):
        """The point-wise losses (boundary data, PINN residual) are REPLICATED data: in a sharded run exactly one
        rank may add them to the exchanged sum, or the objective silently becomes world_size * lossb + lossv.
        owns_point_losses=None: rank 0 of the default process group when `elements` is a shard and torch.distributed
        is initialised, else True."""
        self._dist = None
        if owns_point_losses is None:
            owns_point_losses = True
            if elements is not None:
                try:
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized():
                        owns_point_losses = dist.get_rank() == 0
                except ImportError:
                    pass
        self.owns_point_losses = bool(owns_point_losses)

    def _point_slots(self, slots):
        return tuple(slots) if self.owns_point_losses else ()

    def attach_distributed(self, group=None):
        """Connect this rank's engine with its peers (hpv_peer_export / hpv_peer_connect): from here on
        train() sums gradient and losses over the ranks inside the step's last kernel, and every loss this class
        reads back is the GLOBAL value (so that `loss < tresh` stops all ranks at the same iteration)."""
        import torch.distributed as dist
        from .distributed import connect_peers
        if connect_peers(self.engine, group):
            self._dist = (dist, group)
        return self._dist is not None

    def _losses(self):
        """One evaluation of every loss at the current parameters (no update): dict of floats.  In a sharded run
        the values are summed over the ranks (each rank evaluates its element block; the point losses are added by
        their owner only)."""
        self.engine.loss_and_grad()
        v = self.engine.read_losses()
        if getattr(self, "_dist", None) is not None:
            import torch
            dist, group = self._dist
            backend = dist.get_backend(group)
            t = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64))
            if backend == "nccl":
                t = t.cuda()
            dist.all_reduce(t, group=group)
            v = t.cpu().numpy()
        return self._name_losses(v)

    def _fetch(self, name):
        if name == "train_op_Adam":
            self.engine.train_steps(1, want_history=False)
            return None
        if name == "epsilon":
            return np.array([self.engine.get_params()[1]])
        if name == "lossp":
            return self._lossp_value()
        return self._losses()[name]

    def _lossp_value(self):
        return float("nan")


def _tensor_nodes(X_quad, W_quad):
    """Recover the 1-D rule from the reference's flattened tensor grid (P2D:360-365): p = j*Q + i -> (X[i], X[j])."""
    X_quad, W_quad = np.asarray(X_quad, dtype=np.float64), np.asarray(W_quad, dtype=np.float64)
    Q = int(round(np.sqrt(X_quad.shape[0])))
    if Q * Q != X_quad.shape[0]:
        raise ValueError("X_quad must be the flattened Q x Q tensor grid of the reference driver")
    xi, w = X_quad[:Q, 0].copy(), W_quad[:Q, 0].copy()
    ok = (np.array_equal(X_quad[:, 0], np.tile(xi, Q)) and np.array_equal(X_quad[:, 1], np.repeat(xi, Q))
          and np.array_equal(W_quad[:, 0], np.tile(w, Q)) and np.array_equal(W_quad[:, 1], np.repeat(w, Q)))
    if not ok:
        raise ValueError("X_quad / W_quad are not in the reference's meshgrid layout (P2D:362-365)")
    return xi, w


def _grid_elements(gridx, gridy, N_testfcn):
    NEx, NEy = np.size(N_testfcn[0]), np.size(N_testfcn[1])
    lo = np.array([[gridx[ex], gridy[ey]] for ex in range(NEx) for ey in range(NEy)], dtype=np.float64)
    hi = np.array([[gridx[ex + 1], gridy[ey + 1]] for ex in range(NEx) for ey in range(NEy)], dtype=np.float64)
    nt = np.array([[N_testfcn[0][ex], N_testfcn[1][ey]] for ex in range(NEx) for ey in range(NEy)], dtype=np.int32)
    return lo, hi, nt


# =============================================================================================================
class VPINN_Poisson2D(_VPINNBase):
    """main/Poisson-2D/hp-VPINN-Poisson-2D.py :: VPINN (P2D:27-257)."""
    ACT = "tanh"
    PROBLEM = "poisson2d"

    def __init__(self, X_u_train, u_train, X_f_train, f_train, X_quad, W_quad, U_exact_total, F_exact_total,
                 gridx, gridy, N_testfcn, X_test, u_test, layers, var_form=None, scheme=None, loss_his=None,
                 device=0, seed=1234, elements=None, owns_point_losses=None):
        self.var_form = _from_caller("var_form", var_form, 1)
        self.scheme = _from_caller("scheme", scheme, "VPINNs")
        self.loss_his = _from_caller("loss_his", loss_his, None)
        if self.loss_his is None:
            self.loss_his = []
        self._rng = np.random.RandomState(seed)
        X_u_train, X_f_train, X_test = (np.asarray(a, dtype=np.float64) for a in (X_u_train, X_f_train, X_test))
        self.x, self.y, self.utrain = X_u_train[:, 0:1], X_u_train[:, 1:2], np.asarray(u_train, dtype=np.float64)
        self.xquad, self.yquad, self.wquad = np.asarray(X_quad)[:, 0:1], np.asarray(X_quad)[:, 1:2], np.asarray(W_quad)
        self.xf, self.yf, self.ftrain = X_f_train[:, 0:1], X_f_train[:, 1:2], np.asarray(f_train, dtype=np.float64)
        self.xtest, self.ytest, self.utest = X_test[:, 0:1], X_test[:, 1:2], u_test
        self.Nelementx, self.Nelementy = np.size(N_testfcn[0]), np.size(N_testfcn[1])
        self.Ntestx, self.Ntesty = N_testfcn[0][0], N_testfcn[1][0]
        self.U_ext_total, self.F_ext_total = U_exact_total, F_exact_total
        self.layers = list(layers)
        self.epsilon_value = 0.0
        self.weights, self.biases, self.a = self.initialize_NN(self.layers)

        xi, w = _tensor_nodes(X_quad, W_quad)
        lo, hi, nt = _grid_elements(gridx, gridy, N_testfcn)
        ntx, nty = int(np.max(N_testfcn[0])), int(np.max(N_testfcn[1]))
        F = np.asarray(F_exact_total, dtype=np.float64).reshape(lo.shape[0], nty, ntx)
        if elements is not None:                 # element shard of this rank (multi-GPU)
            lo, hi, nt, F = lo[elements], hi[elements], nt[elements], F[elements]
        self._setup_engine(device, self.layers, xi, w, max(ntx, nty), lo, hi, ntx, nty, F, nt, self.var_form)
        self._init_sharding(elements, owns_point_losses)
        e = self.engine
        e.set_point_loss(0, X_u_train, self.utrain, [1, 0, 0, 0, 0], None, 10.0)          # 10*lossb (P2D:125-128)
        e.set_point_loss(1, X_f_train, self.ftrain, [0, 0, 0, 1, 1], None, 1.0)           # lossp (P2D:123, 187-194)
        if self.scheme == "VPINNs":
            e.configure_training(wv=1.0, point_slots=self._point_slots((0,)), lr=0.001)
        elif self.scheme == "PINNs":
            e.configure_training(wv=0.0, point_slots=self._point_slots((0, 1)), lr=0.001)
        else:
            raise ValueError("scheme must be 'VPINNs' or 'PINNs' (P2D:125-128)")
        self.u_test = _Fetch("u_test")

    def _name_losses(self, v):
        out = {"loss": v[0], "lossb": v[2] / 10.0}
        if self.scheme == "VPINNs":
            out["lossv"] = v[1]
        else:
            out["lossv"] = self.engine.varloss_forward(want_residual=False)
            out["lossp"] = v[3]
        return out

    def _lossp_value(self):
        return self.engine.point_loss_forward(1, self.xf.shape[0])[0]

    def _fetch(self, name):
        if name == "u_test":
            return self.predict()
        return super()._fetch(name)

    def net_u(self, x, y):
        return self.engine.net_u(np.hstack((np.asarray(x).reshape(-1, 1), np.asarray(y).reshape(-1, 1)))).reshape(-1, 1)

    def _derivs(self, x, y):
        return self.engine.net_u(np.hstack((np.asarray(x).reshape(-1, 1), np.asarray(y).reshape(-1, 1))), d1=True, d2=True)

    def net_dxu(self, x, y):
        _, d1, d2 = self._derivs(x, y)
        return d1[:, 0:1], d2[:, 0:1]

    def net_dyu(self, x, y):
        _, d1, d2 = self._derivs(x, y)
        return d1[:, 1:2], d2[:, 1:2]

    def net_f(self, x, y):
        _, _, d2 = self._derivs(x, y)
        return d2[:, 0:1] + d2[:, 1:2]

    def Test_fcnx(self, N_test, x):
        return GJ.Test_fcn(N_test, x)

    def Test_fcny(self, N_test, y):
        return GJ.Test_fcn(N_test, y)

    def train(self, nIter):
        """P2D:233-253: Adam step, then the loss at the updated parameters appended to loss_his; print every 100."""
        start_time = time.time()
        done = 0
        while done < nIter:
            n = min(100 - done % 100, nIter - done)
            # loss after update k is the loss the engine reports BEFORE update k+1
            h = self.engine.train_steps(n)
            tail = self._losses()["loss"]
            after = np.concatenate([h[1:, 0], [tail]])
            self.loss_his.extend(float(v) for v in after)
            if done % 100 == 0:
                elapsed = time.time() - start_time
                print('It: %d, Loss: %.3e, Time: %.2f' % (done, after[0], elapsed))
                start_time = time.time()
            done += n
        self._pull()

    def predict(self):
        return self.net_u(self.xtest, self.ytest)


# =============================================================================================================
class VPINN_Poisson1D(_VPINNBase):
    """main/Poisson-1D/hp-VPINN-Poisson-1D.py :: VPINN (P1D:30-224)."""
    ACT = "sin"
    PROBLEM = "poisson1d"

    def __init__(self, X_u_train, u_train, X_quad, W_quad, F_exact_total, grid, X_test, u_test, layers, X_f_train,
                 f_train, var_form=None, lossb_weight=None, LR=None, total_record=None, device=0, seed=1234,
                 elements=None, owns_point_losses=None):
        self.var_form = _from_caller("var_form", var_form, 1)
        self.lossb_weight = _from_caller("lossb_weight", lossb_weight, 1)
        self.LR = _from_caller("LR", LR, 0.001)
        self.total_record = _from_caller("total_record", total_record, None)
        if self.total_record is None:
            self.total_record = []
        self._rng = np.random.RandomState(seed)
        self.x, self.u = np.asarray(X_u_train, dtype=np.float64), np.asarray(u_train, dtype=np.float64)
        self.xf, self.f = np.asarray(X_f_train, dtype=np.float64), np.asarray(f_train, dtype=np.float64)
        self.xquad, self.wquad = np.asarray(X_quad, dtype=np.float64), np.asarray(W_quad, dtype=np.float64)
        self.xtest, self.utest = np.asarray(X_test, dtype=np.float64), u_test
        self.F_ext_total = F_exact_total
        self.Nelement = np.shape(self.F_ext_total)[0]
        self.N_test = np.shape(self.F_ext_total[0])[0]
        self.layers = list(layers)
        self.epsilon_value = 0.0
        self.weights, self.biases, self.a = self.initialize_NN(self.layers)
        grid = np.asarray(grid, dtype=np.float64)
        lo, hi = grid[:-1, None], grid[1:, None]
        F = np.asarray(F_exact_total, dtype=np.float64).reshape(self.Nelement, 1, self.N_test)
        if elements is not None:
            lo, hi, F = lo[elements], hi[elements], F[elements]
        self._setup_engine(device, self.layers, self.xquad.ravel(), self.wquad.ravel(), self.N_test, lo, hi, self.N_test, 1,
                           F, None, self.var_form)
        self._init_sharding(elements, owns_point_losses)
        e = self.engine
        slots = ()
        if self.lossb_weight != 0:
            e.set_point_loss(0, self.x, self.u, [1, 0, 0, 0, 0], None, float(self.lossb_weight))   # P1D:98-100
            slots = (0,)
        e.configure_training(wv=1.0, point_slots=self._point_slots(slots), lr=float(self.LR))

    def _name_losses(self, v):
        lb = v[2] / self.lossb_weight if self.lossb_weight != 0 else float("nan")
        return {"loss": v[0], "lossv": v[1], "lossb": lb}

    def net_u(self, x):
        return self.engine.net_u(np.asarray(x, dtype=np.float64).reshape(-1, 1)).reshape(-1, 1)

    def net_du(self, x):
        _, d1, d2 = self.engine.net_u(np.asarray(x, dtype=np.float64).reshape(-1, 1), d1=True, d2=True)
        return d1, d2

    def net_f(self, x):
        return -self.net_du(x)[1]

    def predict(self, x):
        return self.net_u(x)

    def predict_subdomain(self, grid):
        # the reference version reads the never-assigned self.utest_total (P1D:189) and cannot run
        raise AttributeError("'VPINN' object has no attribute 'utest_total' (as in the reference, P1D:185-195)")

    def train(self, nIter, tresh):
        """P1D:201-224: Adam step every iteration; at it % 10 == 0 read loss/lossb/lossv after that iteration's
        update, record, stop below tresh; print every 100.  The nine updates in between run back to back on the
        device without a read-back."""
        start_time = time.time()
        it = 0
        loss_valueb = loss_valuev = float("nan")
        while it < nIter:
            # iteration `it` (a multiple of 10): one update, then read the three losses (P1D:208-214)
            self.engine.train_steps(1, want_history=False)
            L = self._losses()
            loss_value, loss_valueb, loss_valuev = L["loss"], L["lossb"], L["lossv"]
            self.total_record.append(np.array([it, loss_value]))
            if loss_value < tresh:
                print('It: %d, Loss: %.3e' % (it, loss_value))
                break
            if it % 100 == 0:
                elapsed = time.time() - start_time
                print('It: %d, Lossb: %.3e, Lossv: %.3e, Time: %.2f' % (it, loss_valueb, loss_valuev, elapsed))
                start_time = time.time()
            n = min(9, nIter - it - 1)
            if n > 0:
                self.engine.train_steps(n, want_history=False)   # iterations it+1 .. it+9: no read-back
            it += 10
        self._pull()


# =============================================================================================================
class VPINN_AdvDiff(_VPINNBase):
    """main/AdvDiff-Identification/hp-VPINN-AdvDiff-Identification.py :: VPINN (ADI:58-341)."""
    ACT = "tanh"
    PROBLEM = "advdiff"

    def __init__(self, XT_u_train, u_train, XT_f_train, XT_quad, W_quad, T_quad, WT_quad, grid_x, grid_t, N_testfcn,
                 XT_test, u_test, layers, lb, ub, var_form=None, V=None, LR=None, device=0, seed=1234, elements=None,
                 owns_point_losses=None):
        self.var_form = _from_caller("var_form", var_form, 0)
        self.V = _from_caller("V", V, 1.0)
        self.LR = _from_caller("LR", LR, 0.001)
        self._rng = np.random.RandomState(seed)
        self.lb, self.ub = lb, ub
        XT_u_train, XT_f_train, XT_test = (np.asarray(a, dtype=np.float64) for a in (XT_u_train, XT_f_train, XT_test))
        self.x, self.t, self.u = XT_u_train[:, 0:1], XT_u_train[:, 1:2], np.asarray(u_train, dtype=np.float64)
        self.x_f, self.t_f = XT_f_train[:, 0:1], XT_f_train[:, 1:2]
        self.xquad, self.tquad, self.wquad = np.asarray(XT_quad)[:, 0:1], np.asarray(XT_quad)[:, 1:2], np.asarray(W_quad)
        self.tquad_1d = np.asarray(T_quad, dtype=np.float64)[:, None]
        self.xquad_1d = self.tquad_1d
        self.wquad_1d = np.asarray(WT_quad, dtype=np.float64)[:, None]
        self.xtest, self.ttest, self.utest = XT_test[:, 0:1], XT_test[:, 1:2], u_test
        self.Nelementx, self.Nelementt = np.size(N_testfcn[0]), np.size(N_testfcn[1])
        self.layers = list(layers)
        self.epsilon_value = 1.0                                  # self.epsilon = tf.Variable(1*ones) (ADI:63)
        self.weights, self.biases, self.a = self.initialize_NN(self.layers)
        xi, w = _tensor_nodes(XT_quad, W_quad)
        lo, hi, nt = _grid_elements(grid_x, grid_t, N_testfcn)
        ntx, ntt = int(np.max(N_testfcn[0])), int(np.max(N_testfcn[1]))
        if elements is not None:
            lo, hi, nt = lo[elements], hi[elements], nt[elements]
        self._setup_engine(device, self.layers, xi, w, max(ntx, ntt), lo, hi, ntx, ntt, None, nt, self.var_form, float(self.V))
        self._init_sharding(elements, owns_point_losses)
        e = self.engine
        e.set_point_loss(0, XT_u_train, self.u, [1, 0, 0, 0, 0], None, 10.0)                 # lossb = 10*mean(.) (ADI:184)
        # strong-form residual u_t + V u_x - eps u_xx against 0 (net_f, ADI:247-253; lossp, ADI:186 -- not in the loss)
        e.set_point_loss(1, XT_f_train, np.zeros(XT_f_train.shape[0]), [0, float(self.V), 1, 0, 0], [0, 0, 0, -1, 0], 1.0)
        e.configure_training(wv=1.0, point_slots=self._point_slots((0,)), train_eps=True, lr=float(self.LR))
        self.u_NN_test = _Fetch("u_NN_test")

    def _name_losses(self, v):
        return {"loss": v[0], "lossv": v[1], "lossb": v[2]}

    def _lossp_value(self):
        return self.engine.point_loss_forward(1, self.x_f.shape[0])[0]

    def _fetch(self, name):
        if name == "u_NN_test":
            return self.net_u(self.xtest, self.ttest)
        return super()._fetch(name)

    def net_u(self, x, t):
        return self.engine.net_u(np.hstack((np.asarray(x).reshape(-1, 1), np.asarray(t).reshape(-1, 1)))).reshape(-1, 1)

    def _derivs(self, x, t):
        return self.engine.net_u(np.hstack((np.asarray(x).reshape(-1, 1), np.asarray(t).reshape(-1, 1))), d1=True, d2=True)

    def net_dxu(self, x, t):
        _, d1, d2 = self._derivs(x, t)
        return d1[:, 0:1], d2[:, 0:1]

    def net_dtu(self, x, t):
        _, d1, _ = self._derivs(x, t)
        return d1[:, 1:2]

    def net_f(self, x, t):
        _, d1, d2 = self._derivs(x, t)
        eps = self.engine.get_params()[1]
        return d1[:, 1:2] + self.V * d1[:, 0:1] - eps * d2[:, 0:1]

    def callback(self, lossv, lossb):
        print('Lossv: %e, Lossb: %e' % (lossv, lossb))

    def predict(self):
        return self.net_u(self.xtest, self.ttest)

    def train(self, nIter, tresh):
        """ADI:291-341, including its return tuple (error_records, total_records, u_records, u_records_iterhis,
        total_time_train); total_time_train accumulates only the train-op time as the reference does."""
        total_time_train, min_loss = 0.0, 1e16
        start_time = time.time()
        total_records, u_records_iterhis = [], []
        u_records = None
        loss_value = float("nan")
        it = 0
        while it < nIter:
            t0 = time.time()
            self.engine.train_steps(1, want_history=False)
            self.engine.sync()
            elapsed_time_train = time.time() - t0
            total_time_train += elapsed_time_train
            L = self._losses()
            loss_value, loss_valueb, loss_valuev = L["loss"], L["lossb"], L["lossv"]
            loss_valuep, a_value = 1, 1
            epsilon_value = self.engine.get_params()[1]
            total_records.append(np.array([it, loss_value, epsilon_value, a_value]))
            if loss_value < tresh:
                print('It: %d, Loss: %.3e' % (it, loss_value))
                break
            if it > 0.9 * nIter and loss_value < min_loss:
                min_loss = loss_value
                u_records = self.net_u(self.xtest, self.ttest)
            if it % 100 == 0:
                elapsed = time.time() - start_time
                print('It: %d, Lossv: %.3e, Lossp: %.3e, Lossb: %.3e, Time: %.2f, TrTime: %.4f, epsilon: %.4f'
                      % (it, loss_valuev, loss_valuep, loss_valueb, elapsed, elapsed_time_train, epsilon_value))
                start_time = time.time()
            n = min(9, nIter - it - 1)
            if n > 0:
                t0 = time.time()
                self.engine.train_steps(n, want_history=False)
                self.engine.sync()
                total_time_train += time.time() - t0
            it += 10
        self._pull()
        error_records = [loss_value, 1]
        return error_records, total_records, u_records, u_records_iterhis, total_time_train
