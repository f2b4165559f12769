#!/usr/bin/env python
"""bench.py -- element-residuals/sec of the hp-VPINN variational-loss path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c2|c3|c4|c5]

A "step" is one full pass of the hot path over the element batch: fused forward (element residuals + lossv),
backward (d lossv / d theta), [N>1: one NCCL all-reduce of loss + gradient], TF1-Adam update -- what one
`sess.run(train_op_Adam)` does in the reference (P2D:242).  value = elements processed by all ranks / time.

Workloads (BASELINE.json configs): c3 = 2-D Poisson 8x8 elements, Q=80x80, 60x60 test functions,
MLP [2,20,20,20,1] (the configuration the metric is quoted on; default at N=1); c4 = 32x32 elements of the
same, block-partitioned over the ranks (default at N>1: strong scaling, total work fixed).

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
LAYERS_2D = [2, 20, 20, 20, 1]


# ---------------------------------------------------------------------------------------------------------------
# synthetic workload (set-up side, float64 numpy; not timed)
# ---------------------------------------------------------------------------------------------------------------
def f_ext_2d(x, y, omegax=2 * np.pi, omegay=2 * np.pi, r1=10):
    """Manufactured right-hand side of the reference driver (P2D:300-310)."""
    gtemp = (-0.1 * (omegax ** 2) * np.sin(omegax * x) - (2 * r1 ** 2) * (np.tanh(r1 * x)) / ((np.cosh(r1 * x)) ** 2)) * np.sin(omegay * y) \
        + (0.1 * np.sin(omegax * x) + np.tanh(r1 * x)) * (-omegay ** 2 * np.sin(omegay * y))
    return gtemp


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
    "c3": dict(desc="C3: 2-D Poisson, 8x8 elements, Q=80x80, N_test=60x60, MLP [2,20,20,20,1], var_form 1", ne=8, Q=80, N=60),
    "c4": dict(desc="C4: 2-D Poisson, 32x32 elements, Q=80x80, N_test=60x60, MLP [2,20,20,20,1], var_form 1", ne=32, Q=80, N=60),
}


def build_workload(name, rank=0, world=1):
    import hpv_b200
    from hpv_b200 import GaussJacobiQuadRule_V3 as GJ
    w = WORKLOADS[name]
    ne, Q, N = w["ne"], w["Q"], w["N"]
    X, W = GJ.GaussLobattoJacobiWeights(Q, 0, 0)
    T = GJ.Test_fcn(N, X)
    D1, D2 = GJ.dTest_fcn(N, X)
    g = np.linspace(-1, 1, ne + 1)
    n_el = ne * ne
    from hpv_b200.distributed import shard_bounds
    e0, e1 = shard_bounds(n_el, rank, world)                            # contiguous block of the (ex, ey) order
    lo = np.array([[g[e // ne], g[e % ne]] for e in range(e0, e1)])
    hi = np.array([[g[e // ne + 1], g[e % ne + 1]] for e in range(e0, e1)])
    A = T * W[None, :]
    F = np.zeros((e1 - e0, N, N))
    for k, e in enumerate(range(e0, e1)):                             # F[k][r] = J sum wx phi_r wy phi_k f (P2D:384-414)
        xe = lo[k, 0] + (hi[k, 0] - lo[k, 0]) / 2 * (X + 1)
        ye = lo[k, 1] + (hi[k, 1] - lo[k, 1]) / 2 * (X + 1)
        jac = (hi[k, 0] - lo[k, 0]) / 2 * (hi[k, 1] - lo[k, 1]) / 2
        F[k] = jac * A @ f_ext_2d(xe[None, :], ye[:, None]) @ A.T
    return dict(name=name, desc=w["desc"], n_el_total=n_el, n_el_local=e1 - e0, Q=Q, N=N, X=X, W=W, T=T, D1=D1, D2=D2,
                lo=lo, hi=hi, F=F, theta=xavier_theta(LAYERS_2D), layers=LAYERS_2D, grid=g, ne=ne)


def make_engine(wl, device):
    import hpv_b200
    eng = hpv_b200.Engine(device)
    eng.set_network(wl["layers"], "tanh")
    eng.set_quadrature(wl["X"], wl["W"])
    eng.set_test_tables(wl["T"], wl["D1"], wl["D2"], None)
    eng.set_form("poisson2d", 1)
    eng.set_elements(wl["lo"], wl["hi"], wl["N"], wl["N"], wl["F"])
    eng.set_params(wl["theta"], 0.0)
    eng.configure_training(wv=1.0, point_slots=(), lr=1e-3)
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
# flop accounting (DESIGN.md "Roofline"): dense MACs x 2, per quadrature point of MLP [2,20,20,20,1], var_form 1
# ---------------------------------------------------------------------------------------------------------------
FLOP_FWD_PT = 5000.0          # value pass 1720 + two tangent passes 1640
FLOP_BWD_PT = 10040.0         # reverse sweep: 2 x (adjoint propagation + weight-gradient products), no recompute
FLOP_BWD_PT_DIR = 6760.0      # the same sweep in the directional mode (value + ONE tangent instead of two): what the
                              # kernel's algorithm needs, so that roofline.frac is a pipe utilisation and does not
                              # take credit for the arithmetic the directional form removed
FLOP_PROJ_EL = 2.688e6        # factorised projection, 2 terms


def cpu_baseline(wl, budget_s=20.0):
    """The restated reference CPU path (oracle, `port`), literal op granularity, forward + backward of whole
    elements of this workload until ~budget_s of CPU time is spent."""
    import torch
    from oracle import hpvpinn_oracle as O
    Ws, bs = O.unpack_theta(wl["theta"], wl["layers"])
    ne, N = wl["ne"], wl["N"]
    X, WX, XY, WXY = O.tensor_quadrature(wl["Q"])
    Ffull = np.zeros((ne, ne, N, N))
    Ffull[0, 0] = wl["F"][0]
    Ntf = [ne * [N], ne * [N]]
    t0 = time.perf_counter()
    done = 0
    while True:
        Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
        bt = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
        loss, _ = O.varloss_2d_literal(Wt, bt, XY, WXY, Ffull, wl["grid"], wl["grid"], Ntf, 1, elements=[(0, 0)])
        torch.autograd.grad(loss, Wt + bt, allow_unused=True)
        done += 1
        el = time.perf_counter() - t0
        if el > budget_s * 0.6 or done >= 4:
            break
    return {"value": done / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d element(s) of %s, forward + backward of the literal torch-float64 restatement (oracle/), %.1f s"
                      % (done, wl["name"].upper(), el)}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (restated, oracle/) on the host cores.  Each step is a
    bounded sample of the workload: the MLP + double autograd of one element plus `m` of the 2*Nty*Ntx
    reduce_sum pairs (and their backward), scaled to a whole element."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import hpvpinn_oracle as O
    name = args.workload if args.workload != "auto" else ("c3" if args.gpus == 1 else "c4")
    wl = build_workload(name, 0, max(1, WORKLOADS[name]["ne"] ** 2))          # one element is enough
    Ws, bs = O.unpack_theta(wl["theta"], wl["layers"])
    Q, N = wl["Q"], wl["N"]
    X, WX, XY, WXY = O.tensor_quadrature(Q)
    g = wl["grid"]
    total_budget = 150.0
    per_step = total_budget / max(1, args.steps + args.warmup)
    xq, yq = XY[:, 0:1], XY[:, 1:2]
    tx = O.Test_fcn(N, xq); d1tx, _ = O.dTest_fcn(N, xq)
    ty = tx; d1ty = d1tx
    w0, w1 = WXY[:, 0:1], WXY[:, 1:2]
    jx = jy = (g[1] - g[0]) / 2
    jac = jx * jy

    def one_step(m):
        """m (k,r) pairs of each of the two var_form-1 terms."""
        Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
        bt = [torch.tensor(b, dtype=torch.float64, requires_grad=True) for b in bs]
        t0 = time.perf_counter()
        x = torch.tensor(g[0] + jx * (xq + 1)).requires_grad_(True)
        y = torch.tensor(g[0] + jy * (yq + 1)).requires_grad_(True)
        u = O.neural_net(torch.cat([x, y], 1), Wt, bt, "tanh")
        d1x, d2x = O.net_d_autograd([x, y], Wt, bt, "tanh", 0)
        d1y, d2y = O.net_d_autograd([x, y], Wt, bt, "tanh", 1)
        t_base = time.perf_counter() - t0
        acc = 0
        cnt = 0
        for k in range(N):
            for r in range(N):
                if cnt >= m:
                    break
                u1 = jac / jx * torch.sum(torch.as_tensor(w0 * d1tx[r] * w1 * ty[k]) * d1x)
                u2 = jac / jy * torch.sum(torch.as_tensor(w0 * tx[r] * w1 * d1ty[k]) * d1y)
                acc = acc + torch.square(-u1 - u2 - float(wl["F"][0][k, r]))
                cnt += 1
            if cnt >= m:
                break
        loss = acc / (N * N)
        torch.autograd.grad(loss, Wt + bt, allow_unused=True)
        t_all = time.perf_counter() - t0
        return t_base, t_all

    m = 4
    tb, ta = one_step(m)                                          # calibrate
    per_pair = max(1e-6, (ta - tb) / m)
    m = int(max(1, min(N * N, (per_step - tb) / per_pair)))
    for _ in range(args.warmup):
        one_step(m)
    t_el = []
    for _ in range(args.steps):
        tb, ta = one_step(m)
        t_el.append(tb + (ta - tb) * (N * N) / m)                 # scaled to one whole element
    t_mean = float(np.mean(t_el))
    value = 1.0 / t_mean
    sample = "per step: MLP + double autograd of one %s element, %d of %d (k,r) pairs per term + backward, scaled to a whole element" % (
        name.upper(), m, N * N)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_mean * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "step": "forward + backward of the variational loss, per element"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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
    evs = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - l0
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    losses = eng.read_losses()

    # ---- forward-only (lossv evaluation), device-timed ----
    evs = []
    barrier()
    for _ in range(max(10, args.steps // 4)):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.forward_async()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    fwd_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    tf = torch.tensor([fwd_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tf, op=dist.ReduceOp.MAX)
    fwd_ms = float(tf.item())

    # ---- end to end through the C ABI with host buffers ----
    Fh = torch.from_numpy(np.ascontiguousarray(wl["F"], dtype=np.float32)).pin_memory()
    theta_host, _ = eng.get_params()
    n_e2e = max(10, min(args.steps, 200))
    for _ in range(3):
        eng.update_rhs_f32(Fh.data_ptr()); eng.set_params(theta_host, 0.0); eng.loss_and_grad()
        if world > 1:
            dist.all_reduce(red)
        eng.read_losses_and_grad()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        eng.update_rhs_f32(Fh.data_ptr())
        eng.set_params(theta_host, 0.0)
        eng.loss_and_grad()
        if world > 1:
            dist.all_reduce(red)
        lv, gv, _ = eng.read_losses_and_grad()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = Fh.numel() * 4 + (eng.n_params + 1) * 8 + (len(theta_host)) * 4
    d2h = 8 * 4 + (eng.n_params + 1) * 8

    if rank == 0:
        n_el = wl["n_el_total"]
        ms_per_step = total_ms / args.steps
        value = n_el * args.steps / (total_ms * 1e-3)
        info = eng.kernel_info()
        npts_local = wl["n_el_local"] * wl["Q"] ** 2
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        flops_bwd = npts_local * (FLOP_BWD_PT_DIR if info.get("bwd_directional") else FLOP_BWD_PT)
        ach = flops_bwd / (kern["mlpbwd"] * 1e-6) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("mlpbwd_dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": "hpv_mlpbwd_kernel (MLP reverse sweep, dominant)", "bound": "fp32-ffma", "achieved": ach,
                    "peak": peak[0], "unit": "TFLOP/s", "frac": ach / peak[0],
                    "peak_source": "FP32 FFMA probe kernel measured in this run (hpv_probe_fp32_peak); MEASURED_PEAKS.json "
                                   "carries only HBM GB/s and bf16 tensor TFLOP/s, neither bounds an fp32 FFMA kernel",
                    "algorithmic_flops_per_launch": flops_bwd, "launch_us": kern["mlpbwd"], "traffic": traffic,
                    "reverse_sweep": "directional (1 tangent)" if info.get("bwd_directional") else "two tangents",
                    "frac_of_measured_bf16_tensor_peak": (ach / peaks["bf16_tflops"]) if "bf16_tflops" in peaks else None,
                    "kernels": {
                        "varfwd": {"us": kern["varfwd"], "algorithmic_gflop": (npts_local * FLOP_FWD_PT + wl["n_el_local"] * FLOP_PROJ_EL) / 1e9},
                        "adjproj": {"us": kern["adjproj"], "algorithmic_gflop": wl["n_el_local"] * FLOP_PROJ_EL / 1e9},
                        "mlpbwd": {"us": kern["mlpbwd"], "algorithmic_gflop": flops_bwd / 1e9},
                        "gradreduce+unpad": {"us": kern["gradreduce+unpad"]}}}
        for kname, kv in roofline["kernels"].items():
            if "algorithmic_gflop" in kv:
                kv["tflops"] = kv["algorithmic_gflop"] * 1e9 / (kv["us"] * 1e-6) / 1e12
                kv["frac_fp32_peak"] = kv["tflops"] / peak[0]
        cpu = cpu_baseline(wl) if not args.no_cpu_baseline else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if name == "c4" else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "elements_total": n_el, "elements_per_gpu": wl["n_el_local"],
                           "step": "fused forward (residuals + lossv) + backward (d theta) + %sAdam" % (
                               ("peer-memory gradient exchange (in the reduction kernel) + " if peer else "NCCL all-reduce + ") if world > 1 else ""),
                           "parallelism": "elements block-partitioned over %d GPU(s)" % world,
                           "l2": "flushed between timed steps (256 MB fill, untimed)", "launch_geometry": info},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": n_el * n_e2e / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "ms_per_step": e2e_s / n_e2e * 1e3, "steps": n_e2e,
                        "what": "per step: hpv_update_rhs_f32 (pinned host F_ext) + hpv_set_params (host theta) + hpv_loss_and_grad "
                                "+ hpv_read_losses_and_grad (host), wall clock"},
                "forward_only": {"value": n_el / (fwd_ms * 1e-3), "unit": UNIT, "ms": fwd_ms},
                "roofline": roofline, "cpu_baseline": cpu, "loss": float(losses[0])}
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
    ap.add_argument("--workload", default="auto", choices=["auto", "c3", "c4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: gradient sum inside the step's last kernel over peer memory (default) or NCCL all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
