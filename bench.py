#!/usr/bin/env python
"""bench.py -- element-residuals/sec of the hp-VPINN variational-loss path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c1|c2|c3|c4|c5]

A "step" is one full pass of the hot path over the element batch: fused forward (element residuals + lossv),
backward (d lossv / d theta), [N>1: sum of loss + gradient over the GPUs], TF1-Adam update -- what one
`sess.run(train_op_Adam)` does in the reference (P2D:242).  value = elements processed by all ranks / time.

Workloads (BASELINE.json configs): c3 = 2-D Poisson 8x8 elements, Q=80x80, 60x60 test functions, MLP [2,20,20,20,1]
(the configuration the metric is quoted on; default at N=1); c4 = 32x32 elements of the same, block-partitioned
over the ranks (default at N>1: strong scaling, total work fixed -- the N=1 line carries the C4 single-GPU step as
`strong_scaling_base`); c2 = 1-D Poisson, 16 elements, Q=80, 60 test functions, [1,20,20,20,1]; c1 = 1-D Poisson, 4 elements, Q=50, 5 test
functions, [1,5,5,1] (the reference's own CPU-sized case); c5 = AdvDiff
identification, 20 elements, Q=80x80, 60x60 test functions, [2,20,20,20,1] + eps.

Timed on the device with CUDA events around every step (L2 flushed by an untimed 256 MB fill between steps),
max over ranks.  `e2e` is the same step driven through the C ABI with HOST buffers: the element batch's
right-hand side (pinned fp32) and the parameters go host->device, loss and gradient come back, every step.
`--impl reference` times the restated reference CPU path (oracle/, torch float64, literal op granularity:
TensorFlow itself cannot be installed here) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "element-residuals/sec (varloss evals/s)"
UNIT = "element-residuals/s"


# ---------------------------------------------------------------------------------------------------------------
# synthetic workloads (set-up side, float64 numpy; not timed)
# ---------------------------------------------------------------------------------------------------------------
def f_ext_2d(x, y, omegax=2 * np.pi, omegay=2 * np.pi, r1=10):
    """Manufactured right-hand side of the reference driver (P2D:300-310)."""
    gtemp = (-0.1 * (omegax ** 2) * np.sin(omegax * x) - (2 * r1 ** 2) * (np.tanh(r1 * x)) / ((np.cosh(r1 * x)) ** 2)) * np.sin(omegay * y) \
        + (0.1 * np.sin(omegax * x) + np.tanh(r1 * x)) * (-omegay ** 2 * np.sin(omegay * y))
    return gtemp


def f_ext_1d(x, omega=8 * np.pi, amp=1, r1=80):
    """Manufactured right-hand side of the reference driver (P1D:248-257): -u'' of u = amp*(0.1 sin(omega x) + tanh(r1 x))."""
    return -amp * (-0.1 * (omega ** 2) * np.sin(omega * x) - (2 * r1 ** 2) * (np.tanh(r1 * x)) / ((np.cosh(r1 * x)) ** 2))


def xavier_theta(layers, seed=1234):
    rng = np.random.default_rng(seed)
    parts = []
    for l in range(len(layers) - 1):
        std = np.sqrt(2.0 / (layers[l] + layers[l + 1]))
        w = rng.standard_normal((layers[l], layers[l + 1]))
        bad = np.abs(w) > 2
        while bad.any():
            w[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(w) > 2
        parts += [(std * w).ravel(), np.zeros(layers[l + 1])]
    return np.concatenate(parts)


WORKLOADS = {
    "c1": dict(kind="poisson1d", ne=4, Q=50, N=5, layers=[1, 5, 5, 1], act="sin", vf=1, nch=3, n_terms=1,
               desc="C1: 1-D Poisson, 4 elements, Q=50, N_test=5, MLP [1,5,5,1], var_form 1 (the reference's own CPU-sized case)"),
    "c2": dict(kind="poisson1d", ne=16, Q=80, N=60, layers=[1, 20, 20, 20, 1], act="sin", vf=1, nch=3, n_terms=1,
               desc="C2: 1-D Poisson, 16 elements, Q=80, N_test=60, MLP [1,20,20,20,1], var_form 1"),
    "c3": dict(kind="poisson2d", ne=8, Q=80, N=60, layers=[2, 20, 20, 20, 1], act="tanh", vf=1, nch=3, n_terms=2,
               desc="C3: 2-D Poisson, 8x8 elements, Q=80x80, N_test=60x60, MLP [2,20,20,20,1], var_form 1"),
    "c4": dict(kind="poisson2d", ne=32, Q=80, N=60, layers=[2, 20, 20, 20, 1], act="tanh", vf=1, nch=3, n_terms=2,
               desc="C4: 2-D Poisson, 32x32 elements, Q=80x80, N_test=60x60, MLP [2,20,20,20,1], var_form 1"),
    "c5": dict(kind="advdiff", ne=20, Q=80, N=60, layers=[2, 20, 20, 20, 1], act="tanh", vf=0, nch=4, n_terms=1,
               desc="C5: AdvDiff identification, 20x1 elements in (x,t), Q=80x80, N_test=60x60, MLP [2,20,20,20,1] + eps, var_form 0"),
}


def build_workload(name, rank=0, world=1):
    import hpv_b200  # noqa: F401
    from hpv_b200 import GaussJacobiQuadRule_V3 as GJ
    from hpv_b200.distributed import shard_bounds
    w = WORKLOADS[name]
    ne, Q, N, kind = w["ne"], w["Q"], w["N"], w["kind"]
    X, W = GJ.GaussLobattoJacobiWeights(Q, 0, 0)
    T = GJ.Test_fcn(N, X)
    D1, D2 = GJ.dTest_fcn(N, X)
    d1b, _ = GJ.dTest_fcn(N, np.array([-1.0, 1.0]))
    A = T * W[None, :]
    wl = dict(name=name, kind=kind, desc=w["desc"], Q=Q, N=N, X=X, W=W, T=T, D1=D1, D2=D2, d1b=d1b, layers=w["layers"], act=w["act"],
              vf=w["vf"], nch=w["nch"], n_terms=w["n_terms"], theta=xavier_theta(w["layers"]), eps=0.0, V=1.0, ne=ne)
    if kind == "poisson2d":
        g = np.linspace(-1, 1, ne + 1)
        n_el = ne * ne
        e0, e1 = shard_bounds(n_el, rank, world)                        # contiguous block of the (ex, ey) order
        lo = np.array([[g[e // ne], g[e % ne]] for e in range(e0, e1)])
        hi = np.array([[g[e // ne + 1], g[e % ne + 1]] for e in range(e0, e1)])
        F = np.zeros((e1 - e0, N, N))
        for k in range(e1 - e0):                                        # F[k][r] = J sum wx phi_r wy phi_k f (P2D:384-414)
            xe = lo[k, 0] + (hi[k, 0] - lo[k, 0]) / 2 * (X + 1)
            ye = lo[k, 1] + (hi[k, 1] - lo[k, 1]) / 2 * (X + 1)
            jac = (hi[k, 0] - lo[k, 0]) / 2 * (hi[k, 1] - lo[k, 1]) / 2
            F[k] = jac * A @ f_ext_2d(xe[None, :], ye[:, None]) @ A.T
        wl.update(grid=g, lo=lo, hi=hi, F=F, ntx=N, nty=N, pts_per_el=Q * Q)
    elif kind == "poisson1d":
        g = np.linspace(-1, 1, ne + 1)
        n_el = ne
        e0, e1 = shard_bounds(n_el, rank, world)
        lo, hi = g[e0:e1, None], g[e0 + 1:e1 + 1, None]
        F = np.zeros((e1 - e0, 1, N))
        for k in range(e1 - e0):                                        # F[i] = J sum w f phi_i (P1D:275-294)
            xe = lo[k, 0] + (hi[k, 0] - lo[k, 0]) / 2 * (X + 1)
            F[k, 0] = (hi[k, 0] - lo[k, 0]) / 2 * A @ f_ext_1d(xe)
        wl.update(grid=g, lo=lo, hi=hi, F=F, ntx=N, nty=1, pts_per_el=Q)
    else:                                                               # advdiff: x in [-1,1], one slab t in [0,1]; RHS = 0 (ADI:180)
        gx, gt = np.linspace(-1, 1, ne + 1), np.array([0.0, 1.0])
        n_el = ne
        e0, e1 = shard_bounds(n_el, rank, world)
        lo = np.array([[gx[e], gt[0]] for e in range(e0, e1)])
        hi = np.array([[gx[e + 1], gt[1]] for e in range(e0, e1)])
        wl.update(grid=gx, grid_t=gt, lo=lo, hi=hi, F=None, ntx=N, nty=N, pts_per_el=Q * Q, eps=1.0)
    wl.update(n_el_total=n_el, n_el_local=e1 - e0)
    return wl


def make_engine(wl, device):
    import hpv_b200
    eng = hpv_b200.Engine(device)
    eng.set_network(wl["layers"], wl["act"])
    eng.set_quadrature(wl["X"], wl["W"])
    eng.set_test_tables(wl["T"], wl["D1"], wl["D2"], wl["d1b"])
    eng.set_form(wl["kind"], wl["vf"], wl["V"])
    eng.set_elements(wl["lo"], wl["hi"], wl["ntx"], wl["nty"], wl["F"])
    eng.set_params(wl["theta"], wl["eps"])
    eng.configure_training(wv=1.0, point_slots=(), lr=1e-3, train_eps=(wl["kind"] == "advdiff"))
    return eng


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------
# flop / byte accounting (DESIGN.md "Roofline"): dense MACs x 2 per quadrature point
#   value pass V = 2 sum in*out, each tangent / second-derivative channel T = V - 2 d_in H1
# ---------------------------------------------------------------------------------------------------------------
def flops_fwd_pt(layers, nch):
    V = 2.0 * sum(layers[i] * layers[i + 1] for i in range(len(layers) - 1))
    return V + (nch - 1) * (V - 2.0 * layers[0] * layers[1])


def flops_bwd_pt(layers, nch):
    """Reverse sweep: adjoint propagation + weight-gradient products of every carried channel (~2x the forward
    products; the forward recompute inside the backward kernel is NOT counted)."""
    return 2.0 * flops_fwd_pt(layers, nch) + 2.0 * layers[0] * layers[1]


def flops_proj_el(wl):
    Q, N = wl["Q"], wl["N"]
    per_term = 2.0 * (N * Q * Q + N * N * Q) if wl["kind"] != "poisson1d" else 2.0 * N * Q
    return wl["n_terms"] * per_term


# ---------------------------------------------------------------------------------------------------------------
# the reference's CPU path, restated (oracle/, `port`): literal op granularity, torch float64 on the host cores
# ---------------------------------------------------------------------------------------------------------------
class LiteralElement2D:
    """One C3/C4-shaped element of P2D:68-120 in the literal form: MLP + double autograd per derivative helper, then
    m of the Nty*Ntx (k, r) reduce_sum pairs of each var_form-1 term and their backward, scaled to the whole element.
    The test-function tables are built ONCE here, outside the timed step (TensorFlow evaluates them at graph
    construction, P2D:85-88)."""

    def __init__(self, wl):
        import torch
        from oracle import hpvpinn_oracle as O
        self.torch, self.O, self.wl = torch, O, wl
        self.Ws, self.bs = O.unpack_theta(wl["theta"], wl["layers"])
        self.N = wl["N"]
        X, WX, XY, WXY = O.tensor_quadrature(wl["Q"])
        self.xq, self.yq = XY[:, 0:1], XY[:, 1:2]
        tx = O.Test_fcn(self.N, self.xq)
        d1tx, _ = O.dTest_fcn(self.N, self.xq)
        ty = O.Test_fcn(self.N, self.yq)
        d1ty, _ = O.dTest_fcn(self.N, self.yq)
        w0, w1 = WXY[:, 0:1], WXY[:, 1:2]
        # numpy pre-multiplication w*test*w*test per (k, r), as the reference's graph constants (P2D:99-105)
        self.c1 = lambda k, r: torch.as_tensor(w0 * d1tx[r] * w1 * ty[k])
        self.c2 = lambda k, r: torch.as_tensor(w0 * tx[r] * w1 * d1ty[k])
        g = wl["grid"]
        self.g0, self.jx, self.jy = g[0], (g[1] - g[0]) / 2, (g[1] - g[0]) / 2
        self.F0 = wl["F"][0]

    def step(self, m):
        torch, O, N = self.torch, self.O, self.N
        Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in self.Ws]
        bt = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in self.bs]
        t0 = time.perf_counter()
        x = torch.tensor(self.g0 + self.jx * (self.xq + 1)).requires_grad_(True)
        y = torch.tensor(self.g0 + self.jy * (self.yq + 1)).requires_grad_(True)
        O.neural_net(torch.cat([x, y], 1), Wt, bt, "tanh")
        d1x, _ = O.net_d_autograd([x, y], Wt, bt, "tanh", 0)
        d1y, _ = O.net_d_autograd([x, y], Wt, bt, "tanh", 1)
        t_base = time.perf_counter() - t0
        jac = self.jx * self.jy
        acc, cnt = 0, 0
        for k in range(N):
            for r in range(N):
                if cnt >= m:
                    break
                u1 = jac / self.jx * torch.sum(self.c1(k, r) * d1x)
                u2 = jac / self.jy * torch.sum(self.c2(k, r) * d1y)
                acc = acc + torch.square(-u1 - u2 - float(self.F0[k, r]))
                cnt += 1
            if cnt >= m:
                break
        torch.autograd.grad(acc / (N * N), Wt + bt, allow_unused=True)
        t_all = time.perf_counter() - t0
        return t_base + (t_all - t_base) * (N * N) / m, t_base, t_all       # seconds per whole element

    def calibrate(self, per_step_s):
        """Pairs per step from WARM calls (the first call pays allocator and thread-pool start-up), at least 512."""
        self.step(64)
        _, tb, ta = self.step(256)
        per_pair = max(1e-6, (ta - tb) / 256)
        return int(max(512, min(self.N * self.N, (per_step_s - tb) / per_pair)))


def factorised_baseline(wl, budget_s=6.0):
    """`ref_factorised` (SURVEY 8d / BASELINE.md 2): the same mathematics in the sum-factorised float64 form
    (tables pre-built; torch BLAS on the host cores), forward + backward of whole elements -- the stronger CPU number."""
    import torch
    from oracle import hpvpinn_oracle as O
    if wl["kind"] != "poisson2d":
        return None
    Ws, bs = O.unpack_theta(wl["theta"], wl["layers"])
    g = wl["grid"]
    N, X, W = wl["N"], wl["X"], wl["W"]
    ne_s = 2                                                             # a 2x2 block of elements per evaluation
    gs = g[:ne_s + 1]
    F = O.rhs_2d_factorised(gs, gs, N, N, X, W)
    fn = lambda Wt, bt: O.varloss_2d_factorised(Wt, bt, X, W, F, gs, gs, N, N, 1)[0]
    O.loss_and_grad(fn, Ws, bs)                                          # warm
    t0 = time.perf_counter()
    done = 0
    while True:
        O.loss_and_grad(fn, Ws, bs)
        done += ne_s * ne_s
        el = time.perf_counter() - t0
        if el > budget_s:
            break
    return {"value": done / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d %s elements, forward + backward of the sum-factorised float64 form (oracle/varloss_2d_factorised), %.1f s" % (
                done, wl["name"].upper(), el)}


def cpu_baseline(wl, budget_s=18.0):
    """The restated reference CPU path (`port`) on a bounded sample of this workload, with the factorised form beside it."""
    import torch
    from oracle import hpvpinn_oracle as O
    if wl["kind"] == "poisson2d":
        lit = LiteralElement2D(wl)
        m = lit.calibrate(budget_s / 3.0)
        ts = [lit.step(m)[0] for _ in range(2)]
        t_el = float(np.mean(ts))
        out = {"value": 1.0 / t_el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": "2 steps: MLP + double autograd of one %s element, %d of %d (k,r) pairs per term + backward, scaled to a whole "
                         "element (tables pre-built, literal torch-float64 restatement, oracle/)" % (wl["name"].upper(), m, wl["N"] ** 2)}
        out["ref_factorised"] = factorised_baseline(wl)
        return out
    Ws, bs = O.unpack_theta(wl["theta"], wl["layers"])
    t0 = time.perf_counter()
    if wl["kind"] == "poisson1d":
        F = wl["F"].reshape(wl["n_el_local"], wl["N"], 1)
        done, reps = 0, 0
        while time.perf_counter() - t0 < budget_s * 0.5 or reps < 2:
            Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
            bt = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
            loss, _ = O.varloss_1d_literal(Wt, bt, wl["X"][:, None], wl["W"][:, None], F, wl["grid"], 1)
            torch.autograd.grad(loss, Wt + bt, allow_unused=True)
            done += wl["n_el_local"]; reps += 1
        el = time.perf_counter() - t0
        sample = "%d passes over the %d elements of %s, literal forward + backward (incl. the per-element table evaluation)" % (reps, wl["n_el_local"], wl["name"].upper())
    else:
        X, WX, XT, WXT = O.tensor_quadrature(wl["Q"])
        XT = XT.copy(); XT[:, 1] = XT[:, 1]                               # reference coordinates; the element map is applied inside
        Ntf = [[wl["N"]] * wl["ne"], [wl["N"]]]
        Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
        bt = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
        eps = torch.tensor([wl["eps"]], dtype=torch.float64, requires_grad=True)
        loss, _ = O.varloss_adi_literal(Wt, bt, eps, XT, WXT, wl["grid"], wl["grid_t"], Ntf, 0, wl["V"], elements=[(0, 0)])
        torch.autograd.grad(loss, Wt + bt + [eps], allow_unused=True)
        done = 1
        el = time.perf_counter() - t0
        sample = "1 element of C5, literal forward + backward (incl. the per-element table evaluation), %.1f s" % el
    return {"value": done / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (restated, oracle/) on the host cores, rank 0 only.  Each
    step is a bounded sample of the workload (see LiteralElement2D), scaled to a whole element."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    name = args.workload if args.workload != "auto" else ("c3" if args.gpus == 1 else "c4")
    nel = WORKLOADS[name]["ne"] ** (2 if WORKLOADS[name]["kind"] == "poisson2d" else 1)
    wl = build_workload(name, 0, max(1, nel))                           # one element is enough
    if wl["kind"] == "poisson2d":
        lit = LiteralElement2D(wl)
        total_budget = 150.0
        m = lit.calibrate(total_budget / max(1, args.steps + args.warmup))
        for _ in range(args.warmup):
            lit.step(m)
        t_el = [lit.step(m)[0] for _ in range(args.steps)]
        t_mean = float(np.mean(t_el))
        sample = ("per step: MLP + double autograd of one %s element, %d of %d (k,r) pairs per term + backward, scaled to a whole element "
                  "(tables pre-built)" % (name.upper(), m, wl["N"] ** 2))
        value = 1.0 / t_mean
    else:
        cb = cpu_baseline(wl, budget_s=60.0)
        value, sample, t_mean = cb["value"], cb["sample"], 1.0 / cb["value"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_mean * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "step": "forward + backward of the variational loss, per element"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def time_steps(torch, eng, stream, flush, step, n, barrier):
    evs = []
    barrier()
    for _ in range(n):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    return sum(a.elapsed_time(b) for a, b in evs)


def multi_gpu_parity(torch, dist, eng, wl, name, rank, world, local, red, peer, step, stream):
    """N>1: the sharded step against ONE engine holding the whole batch (rank 0), before anything is timed.
    (a) gradient: local loss_and_grad + all-reduce of the reduce buffer vs the single engine's gradient;
    (b) the shipped step (peer exchange or NCCL, + Adam), 3 steps: loss history vs the single engine's, and the
        parameters of all ranks bitwise equal."""
    out = {}
    ref = None
    if rank == 0:
        full = build_workload(name, 0, 1)
        e1 = make_engine(full, local)
        e1.set_stream(stream.cuda_stream)
        e1.loss_and_grad()
        l_ref, g_ref, _ = e1.read_losses_and_grad()
        h_ref = e1.train_steps(3)
        t_ref, _ = e1.get_params()
        e1.close()
        ref = (l_ref[0], g_ref, h_ref[:, 0].copy(), t_ref)
    eng.loss_and_grad()
    dist.all_reduce(red)
    l, g, _ = eng.read_losses_and_grad()
    h = eng.train_steps(3) if (peer or world == 1) else None
    if h is None:
        hs = []
        for _ in range(3):
            eng.loss_and_grad(); dist.all_reduce(red); hs.append(eng.read_losses()[0]); eng.adam_step()
        h = np.array(hs)[:, None]
    theta, _ = eng.get_params()
    tt = torch.from_numpy(theta.copy()).cuda()
    gathered = [torch.empty_like(tt) for _ in range(world)]
    dist.all_gather(gathered, tt)
    if rank == 0:
        same = all(bool(torch.equal(gathered[0], x)) for x in gathered[1:])
        out = {"loss_rel": float(abs(l[0] - ref[0]) / abs(ref[0])),
               "grad_rel": float(np.abs(g - ref[1]).max() / np.abs(ref[1]).max()),
               "loss_history_rel_3_steps": float(np.abs(h[:, 0] - ref[2]).max() / np.abs(ref[2]).max()),
               "theta_rel_after_3_steps": float(np.abs(theta - ref[3]).max() / max(1.0, np.abs(ref[3]).max())),
               "ranks_bitwise_equal": bool(same),
               "against": "one engine holding all %d elements on rank 0's GPU, same parameters" % wl["n_el_total"]}
    # back to the initial state for the timed part
    eng.set_params(wl["theta"], wl["eps"])
    eng.reset_optimizer()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    name = args.workload if args.workload != "auto" else ("c3" if world == 1 else "c4")
    wl = build_workload(name, rank, world)
    eng = make_engine(wl, local)
    # a dedicated (non-default) torch stream carries the engine's launches, NCCL and the timing events, so
    # that the events bracket exactly the step's work (hpv_set_stream(NULL) would mean the engine's own stream)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    from hpv_b200.distributed import reduce_tensor, connect_peers
    red = reduce_tensor(eng) if world > 1 else None
    peer = world > 1 and args.collective == "peer" and connect_peers(eng)

    def step():
        if world == 1 or peer:
            # forward, adjoint projection, reverse sweep, reduction [+ peer-memory exchange over NVLink] + Adam
            eng.train_steps(1, want_history=False)
        else:
            eng.loss_and_grad()
            dist.all_reduce(red)
            eng.adam_step()

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = None
    if world > 1:
        parity = multi_gpu_parity(torch, dist, eng, wl, name, rank, world, local, red, peer, step, stream)
        barrier()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- kernel-only numbers (rank 0): per-kernel device time of the step's launches ----
    kern = {}
    if rank == 0:
        for what, nm in ((0, "varfwd"), (1, "adjproj"), (2, "mlpbwd"), (3, "gradreduce+unpad")):
            kern[nm] = eng.time_kernel(what, 30)
        peak = {v: eng.probe_fp32_peak(v) for v in (0,)}
    barrier()

    # ---- device-timed steps ----
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    total_ms = time_steps(torch, eng, stream, flush, step, args.steps, barrier)
    clocks = sampler.stop()
    launches = eng.launch_count() - l0
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    losses = eng.read_losses()

    # ---- forward-only (lossv evaluation), device-timed ----
    nf = max(10, args.steps // 4)
    fwd_ms = time_steps(torch, eng, stream, flush, eng.forward_async, nf, barrier) / nf
    tf = torch.tensor([fwd_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    fwd_ms = float(tf.item())

    # ---- end to end through the C ABI with host buffers (the shipped step: at N>1 the exchange the timed steps use) ----
    has_F = wl["F"] is not None
    Fh = torch.from_numpy(np.ascontiguousarray(wl["F"], dtype=np.float32)).pin_memory() if has_F else None
    theta_host, eps_host = eng.get_params()

    def e2e_step():
        if has_F:
            eng.update_rhs_f32(Fh.data_ptr())
        eng.set_params(theta_host, eps_host)
        if world > 1 and peer:
            eng.train_steps(1, want_history=False)            # gradient + losses summed over the GPUs inside the last kernel
        else:
            eng.loss_and_grad()
            if world > 1:
                dist.all_reduce(red)
        return eng.read_losses_and_grad()

    n_e2e = max(10, min(args.steps, 200))
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        lv, gv, _ = e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = (Fh.numel() * 4 if has_F else 0) + (eng.n_params + 1) * 8 + (len(theta_host)) * 4
    d2h = 8 * 4 + (eng.n_params + 1) * 8

    # ---- N = 1 default run: the C4 single-GPU step, the base of the strong-scaling series the N > 1 lines report ----
    base = None
    if world == 1 and name == "c3" and args.workload == "auto" and not args.no_scaling_base:
        eng.set_params(wl["theta"], wl["eps"]); eng.reset_optimizer()
        w4 = build_workload("c4", 0, 1)
        e4 = make_engine(w4, local)
        e4.set_stream(stream.cuda_stream)
        s4 = lambda: e4.train_steps(1, want_history=False)
        for _ in range(3):
            s4()
        n4 = max(10, min(60, args.steps // 10))
        ms4 = time_steps(torch, e4, stream, flush, s4, n4, barrier) / n4
        base = {"workload": w4["desc"], "ms_per_step": ms4, "value": w4["n_el_total"] / (ms4 * 1e-3), "unit": UNIT, "steps": n4,
                "loss": float(e4.read_losses()[0]), "note": "same step, all 1024 elements on this one GPU: divide the N>1 lines' value by this"}
        e4.close()

    if rank == 0:
        n_el = wl["n_el_total"]
        ms_per_step = total_ms / args.steps
        value = n_el * args.steps / (total_ms * 1e-3)
        info = eng.kernel_info()
        npts_local = wl["n_el_local"] * wl["pts_per_el"]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        nch_bwd = 2 if info.get("bwd_directional") else wl["nch"]
        f_fwd = npts_local * flops_fwd_pt(wl["layers"], wl["nch"]) + wl["n_el_local"] * flops_proj_el(wl)
        f_adj = wl["n_el_local"] * flops_proj_el(wl)
        f_bwd = npts_local * flops_bwd_pt(wl["layers"], nch_bwd)
        ach = f_bwd / (kern["mlpbwd"] * 1e-6) / 1e12
        # DRAM bytes of the dominant kernel: measured by ncu per (workload, GPU count) where a capture exists
        # (profiles/traffic.json), next to the algorithmic figure (Gbar read once: terms x points x 4 B)
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get("%s@%d" % (name, world))
            if ent:
                traffic, traffic_src = ent.get("mlpbwd_dram_bytes_per_launch"), ent.get("source")
        except Exception:
            pass
        roofline = {"kernel": "hpv_mlpbwd_kernel (MLP reverse sweep, dominant)", "bound": "fp32-ffma", "achieved": ach,
                    "peak": peak[0], "unit": "TFLOP/s", "frac": ach / peak[0],
                    "peak_source": "FP32 FFMA probe kernel measured in this run (hpv_probe_fp32_peak); MEASURED_PEAKS.json "
                                   "carries only HBM GB/s and bf16 tensor TFLOP/s, neither bounds an fp32 FFMA kernel",
                    "algorithmic_flops_per_launch": f_bwd, "launch_us": kern["mlpbwd"], "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_algorithmic": float(wl["n_terms"] * npts_local * 4),
                    "reverse_sweep": "directional (1 tangent)" if info.get("bwd_directional") else "%d channels" % wl["nch"],
                    "frac_of_measured_bf16_tensor_peak": (ach / peaks["bf16_tflops"]) if "bf16_tflops" in peaks else None,
                    "forward_kernel": "tensor-core form: tcgen05 kind::tf32 3-term split, A in TMEM, TMA-staged tables" if info.get("fwd_tensor_core")
                                      else "FP32-FFMA form",
                    "kernels": {
                        "varfwd": {"us": kern["varfwd"], "algorithmic_gflop": f_fwd / 1e9},
                        "adjproj": {"us": kern["adjproj"], "algorithmic_gflop": f_adj / 1e9},
                        "mlpbwd": {"us": kern["mlpbwd"], "algorithmic_gflop": f_bwd / 1e9},
                        "gradreduce+unpad": {"us": kern["gradreduce+unpad"]}}}
        for kname, kv in roofline["kernels"].items():
            if "algorithmic_gflop" in kv:
                kv["tflops"] = kv["algorithmic_gflop"] * 1e9 / (kv["us"] * 1e-6) / 1e12
                kv["frac_fp32_peak"] = kv["tflops"] / peak[0]
        cpu = cpu_baseline(wl) if not args.no_cpu_baseline else None
        xch = ("peer-memory gradient exchange (in the reduction kernel) + " if peer else "NCCL all-reduce + ") if world > 1 else ""
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "elements_total": n_el, "elements_per_gpu": wl["n_el_local"],
                           "step": "fused forward (residuals + lossv) + backward (d theta) + %sAdam" % xch,
                           "parallelism": "elements block-partitioned over %d GPU(s)" % world,
                           "scaling_series": "strong scaling is reported on C4 (1024 elements, total work fixed): the N=1 line's "
                                             "`strong_scaling_base` is the C4 step on one GPU, its `value` is the C3 headline",
                           "l2": "flushed between timed steps (256 MB fill, untimed)", "launch_geometry": info},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": n_el * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
                        "what": "per step: hpv_update_rhs_f32 (pinned host F_ext) + hpv_set_params (host theta) + %s + hpv_read_losses_and_grad "
                                "(host), wall clock" % ("hpv_train_steps(1) with the in-kernel peer exchange" if (world > 1 and peer) else
                                                        "hpv_loss_and_grad" + (" + NCCL all-reduce" if world > 1 else ""))},
                "forward_only": {"value": n_el / (fwd_ms * 1e-3), "unit": UNIT, "ms": fwd_ms},
                "roofline": roofline, "cpu_baseline": cpu, "loss": float(losses[0])}
        if base is not None:
            line["strong_scaling_base"] = base
        if parity is not None:
            line["parity"] = parity
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true", help="N=1: skip the C4 single-GPU step (strong_scaling_base)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: gradient sum inside the step's last kernel over peer memory (default) or NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
