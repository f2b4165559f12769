"""Runs a few full training steps of a bench workload (for ncu captures; no timing claims are made here)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--steps", type=int, default=4)
a = ap.parse_args()
wl = bench.build_workload(a.workload)
eng = bench.make_engine(wl, 0)
for _ in range(a.steps):
    eng.train_steps(1, want_history=False)
eng.sync()
print("losses", eng.read_losses()[:2], "launches", eng.launch_count())
