"""Kernel time as a function of the number of elements (development aid): separates per-launch overhead and
quantisation from the per-element cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for ne in (2, 3, 4, 6, 8, 10, 12, 16):
    bench.WORKLOADS["x"] = dict(bench.WORKLOADS["c3"], desc="x", ne=ne)
    wl = bench.build_workload("x")
    eng = bench.make_engine(wl, 0)
    t = [eng.time_kernel(w, 30) for w in (0, 1, 2)]
    info = eng.kernel_info()
    print("n_el %4d  fwd %7.1f us (%5.2f/el)  adj %6.1f  bwd %7.1f us (%5.2f/el)  fwd_grid %d bwd_grid %d x %d fwd_tc %d" % (
        ne * ne, t[0], t[0] / ne / ne, t[1], t[2], t[2] / ne / ne, info["fwd_grid"], info["bwd_grid"], info["bwd_block"],
        info.get("fwd_tensor_core", 0)), flush=True)
    eng.close()
