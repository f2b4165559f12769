"""Soak check (development aid): 20 000 graph-replayed training steps of C3, twice from the same start; the two runs must
end with bitwise identical parameters and loss histories, and the loss must have gone down."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

wl = bench.build_workload("c3")
out = []
for run in range(2):
    eng = bench.make_engine(wl, 0)
    t0 = time.perf_counter()
    hist = [eng.train_steps(1000) for _ in range(20)]
    theta = eng.get_params() if hasattr(eng, "get_params") else None
    dt = time.perf_counter() - t0
    h = np.concatenate(hist)
    out.append((h, theta))
    print("run %d: 20000 steps in %.2f s = %.1f us/step (host loop included); loss %.6e -> %.6e" % (run, dt, dt / 20000 * 1e6, h[0, 0], h[-1, 0]), flush=True)
    eng.close()
print("histories bitwise equal:", np.array_equal(out[0][0], out[1][0]))
if out[0][1] is not None:
    pa, pb = out[0][1], out[1][1]
    pa = pa if isinstance(pa, (tuple, list)) else (pa,)
    pb = pb if isinstance(pb, (tuple, list)) else (pb,)
    print("parameters bitwise equal:", all(np.array_equal(np.asarray(a), np.asarray(b)) for a, b in zip(pa, pb)))
assert np.array_equal(out[0][0], out[1][0]) and out[0][0][-1, 0] < out[0][0][0, 0]
