import sys, os, time
sys.path.insert(0, "/root/repo")
import numpy as np
from oracle import hpvpinn_oracle as O
from tests import _cases as C
from hpv_b200.poisson2d import VPINN
d = dict(np.load(os.path.join(C.GOLDEN, "driver_p2d.npz")))
rng = np.random.default_rng(0)
X_f = 2 * rng.random((100, 2)) - 1
f = O.f_ext_2d(X_f[:, 0:1], X_f[:, 1:2])
xs = np.linspace(-1, 1, 41)
X_test = np.array([[a, b] for a in xs for b in xs]); u_test = O.u_ext_2d(X_test[:, 0:1], X_test[:, 1:2])
his = []
m = VPINN(d["X_u_train"], d["u_train"], X_f, f, d["XY_quad"], d["WXY_quad"], None, d["F_ext_total"], d["grid_x"], d["grid_y"],
          [list(d["N_test_x"]), list(d["N_test_y"])], X_test, u_test, [int(v) for v in d["layers"]], var_form=1, scheme="VPINNs", loss_his=his)
import io, contextlib
t0 = time.time()
for k in range(10):
    with contextlib.redirect_stdout(io.StringIO()):
        m.train(1000)
    L = m._losses()
    e = np.linalg.norm(m.predict() - u_test) / np.linalg.norm(u_test)
    print(len(his), "loss %.4e lossb %.4e lossv %.4e relL2 %.3f maxerr %.3f  %.1fs" % (L["loss"], L["lossb"], L["lossv"], e, np.abs(m.predict() - u_test).max(), time.time() - t0), flush=True)
