"""Host-side breakdown of the e2e step of bench.py (development aid): wall time spent inside each of the four C-ABI
calls of a host-driven step, and the step total, C3."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = bench.build_workload("c3")
eng = bench.make_engine(wl, 0)
Fh = torch.from_numpy(np.ascontiguousarray(wl["F"], dtype=np.float32)).pin_memory()
theta, eps = wl["theta"].copy(), 0.0
def step(ts):
    t0 = time.perf_counter(); eng.update_rhs_f32(Fh.data_ptr())
    t1 = time.perf_counter(); eng.set_params(theta, eps)
    t2 = time.perf_counter(); eng.loss_and_grad()
    t3 = time.perf_counter(); out = eng.read_losses_and_grad()
    t4 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0))
    return out
for _ in range(20): step([])
ts = []
for _ in range(300): step(ts)
a = np.array(ts) * 1e6
print("us per call (median / mean): update_rhs %.1f / %.1f | set_params %.1f / %.1f | loss_and_grad %.1f / %.1f | read_losses_and_grad %.1f / %.1f | step %.1f / %.1f" % tuple(
    v for i in range(5) for v in (np.median(a[:, i]), a[:, i].mean())))
dev = eng.time_kernel(0, 20) + eng.time_kernel(1, 20) + eng.time_kernel(2, 20) + eng.time_kernel(3, 20)
print("device kernels (warm L2, back to back): %.1f us" % dev)
