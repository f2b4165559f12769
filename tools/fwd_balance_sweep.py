"""Forward-kernel time at C3 and at 256 elements for a grid of partition-cost parameters (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for ne in (8, 16):
    bench.WORKLOADS["x"] = dict(bench.WORKLOADS["c3"], desc="x", ne=ne)
    wl = bench.build_workload("x")
    for cross in ("1.6", "2.4", "3.2", "4.0"):
        row = []
        for wave2 in ("1.08", "1.16", "1.24"):
            os.environ["HPV_FWD_BALANCE_CROSS"] = cross; os.environ["HPV_FWD_BALANCE_WAVE2"] = wave2
            eng = bench.make_engine(wl, 0)
            t = min(eng.time_kernel(0, 30) for _ in range(3))
            row.append("%.1f" % t)
            eng.close()
        print("n_el %3d cross %s  wave2 1.08/1.16/1.24: %s" % (ne * ne, cross, " ".join(row)), flush=True)
