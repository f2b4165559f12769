"""Development probe for the first GPU call: parity on every golden case, FP32 probes, C3 timing."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import _cases as C, _gpu as G
from oracle import hpvpinn_oracle as O
import hpv_b200

out = {}
for name in C.case_names():
    c = C.load(name); inp = C.engine_inputs(c)
    try:
        eng = G.make_engine(inp)
        loss, res = eng.varloss_forward()
        g, ge = eng.varloss_backward()
        o = C.oracle_lossv(c)
        ores = o[1].reshape(res.shape)
        print("%-14s loss rel %.2e | res rel %.2e | grad rel %.2e" % (name, abs(loss - o[0]) / abs(o[0]),
              np.abs(res - ores).max() / np.abs(ores).max(), np.abs(g - o[2]).max() / np.abs(o[2]).max()), flush=True)
        eng.close()
    except Exception as e:
        print(name, "FAILED", e, flush=True)

eng = hpv_b200.Engine(0)
for v in (0, 1, 2):
    try:
        print("fp32 probe variant", v, "%.1f TFLOP/s" % eng.probe_fp32_peak(v), flush=True)
    except Exception as e:
        print("probe", v, "failed", e)
# C3: 8x8 elements, Q=80, N=60, [2,20,20,20,1]
Q, N = 80, 60
X, W = O.GaussLobattoJacobiWeights(Q, 0, 0)
gx = np.linspace(-1, 1, 9)
lo = np.array([[gx[i], gx[j]] for i in range(8) for j in range(8)]); hi = lo + (gx[1] - gx[0])
F = O.rhs_2d_factorised(gx, gx, N, N, X, W).reshape(64, N, N)
layers = [2, 20, 20, 20, 1]
Ws, bs = O.xavier_params(layers, 1234)
theta = O.pack_theta(Ws, bs)
inp = dict(problem="poisson2d", var_form=1, layers=layers, act="tanh", theta=theta, xi=X, w=W, T=O.Test_fcn(N, X),
           D1=O.dTest_fcn(N, X)[0], D2=O.dTest_fcn(N, X)[1], d1b=None, lo=lo, hi=hi, ntx=N, nty=N, F=F)
e3 = G.make_engine(inp)
t0 = time.time(); loss, res = e3.varloss_forward(); t1 = time.time()
print("C3 first forward %.3fs loss %.10e" % (t1 - t0, loss), flush=True)
g, _ = e3.varloss_backward()
print("C3 info", e3.kernel_info(), flush=True)
lo_, res_o = O.varloss_2d_factorised(Ws, bs, X, W, F.reshape(8, 8, N, N), gx, gx, N, N, 1)
lo_ = float(lo_)
print("C3 oracle loss %.10e rel %.2e res rel %.2e" % (lo_, abs(loss - lo_) / lo_, np.abs(res - res_o.numpy()).max() / np.abs(res_o.numpy()).max()), flush=True)
for what, nm in ((0, "forward"), (1, "adjproj"), (2, "mlpbwd"), (3, "reduce+unpad")):
    print("C3 %-12s %.1f us" % (nm, e3.time_kernel(what, 50)), flush=True)
e3.configure_training(wv=1.0)
h = e3.train_steps(20)
e3.sync(); t0 = time.time(); h = e3.train_steps(200); t1 = time.time()
print("C3 train step %.1f us/step (wall, 200 steps) loss %.4e -> %.4e" % ((t1 - t0) / 200 * 1e6, h[0, 0], h[-1, 0]), flush=True)
