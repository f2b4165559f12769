"""Residual accuracy of the two forward kernels (tensor-core 3xTF32 vs FP32 FFMA) against the float64 oracle at C3 size,
random-init and near-converged parameters (development aid; run with HPV_FWD_TC=0/1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hpvpinn_oracle as O
from tests import _cases as C, _gpu as G
from tests.test_gpu_parity_regimes import _poisson2d_inputs

fx = dict(np.load(os.path.join(C.GOLDEN, "c3_converged.npz")))
layers = [int(v) for v in fx["layers"]]
g = np.linspace(-1, 1, 9)
Ws0, bs0 = O.xavier_params(layers, 1234)
for tag, theta in (("random-init", O.pack_theta(Ws0, bs0)), ("near-converged", fx["theta"])):
    inp, X, W, F = _poisson2d_inputs(g, g, 80, 60, theta, layers, 1)
    Ws, bs = O.unpack_theta(theta, layers)
    l_ref, r_ref = O.varloss_2d_factorised(Ws, bs, X, W, F, g, g, 60, 60, 1)
    r_ref = r_ref.numpy().reshape(64, 60, 60)
    U_ref = r_ref + F.reshape(64, 60, 60)
    eng = G.make_engine(inp)
    loss, res, el = eng.varloss_forward(want_residual=True, want_el_loss=True)
    info = eng.kernel_info()
    err = np.abs(res - r_ref)
    el_ref = (r_ref ** 2).mean(axis=(1, 2))
    print("%s fwd_tc=%d: lossv rel %.2e | max|Res err| %.3e = %.2e of max|U| (%.3f), %.2e of max|Res| | el_loss rel(max) %.2e | worst element %d" % (
        tag, info["fwd_tensor_core"], abs(loss - float(l_ref)) / float(l_ref), err.max(), err.max() / np.abs(U_ref).max(), np.abs(U_ref).max(),
        err.max() / np.abs(r_ref).max(), np.abs(el - el_ref).max() / el_ref.max(), int(err.reshape(64, -1).max(1).argmax())), flush=True)
    # per-point accuracy of the network itself
    pts = np.random.default_rng(0).uniform(-1, 1, (4096, 2))
    u, d1, d2 = eng.net_u(pts, d1=True, d2=True)
    import torch
    uu, dd1, dd2 = O.mlp_forward_mode(pts, [torch.as_tensor(w) for w in Ws], [torch.as_tensor(b) for b in bs], "tanh")
    print("   net_u (FFMA points kernel) max err u %.2e d1 %.2e | max|u| %.2f max|d1| %.2f" % (np.abs(u - uu.numpy().ravel()).max(), np.abs(d1 - dd1.numpy()).max(),
          np.abs(uu.numpy()).max(), np.abs(dd1.numpy()).max()))
    eng.close()
