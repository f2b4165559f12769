"""Timing experiment (development aid): per-CTA / per-warp timeline of the FFMA reverse sweep, from a library built with
HPV_NVCC_EXTRA=-DHPV_EXP_STAMPS (HPV_LIB points to it)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

NST = 24
lib = ctypes.CDLL(os.environ["HPV_LIB"])
lib.hpv_exp_read_bstamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
for name in ("c3", "c4"):
    wl = bench.build_workload(name)
    eng = bench.make_engine(wl, 0)
    t = eng.time_kernel(2, 10)
    info = eng.kernel_info()
    n, nw = info["bwd_grid"], info["bwd_block"] // 32
    buf = np.zeros((n, NST), dtype=np.uint64)
    assert lib.hpv_exp_read_bstamps(buf.ctypes.data, n) == 0
    s = buf.astype(np.float64) * 1e-3
    t0 = s[:, 0].min()
    ends = s[:, 2:2 + nw] - t0
    def rng(v): return "min %7.1f med %7.1f max %7.1f" % (v.min(), np.median(v), v.max())
    print("%s  bwd %.1f us  grid %d x %d warps" % (name, t, n, nw))
    print("   CTA entry after first CTA     ", rng(s[:, 0] - t0))
    print("   prologue + wait               ", rng(s[:, 1] - s[:, 0]))
    print("   warp sweep end (all warps)    ", rng(ends))
    print("   per CTA: first warp to finish ", rng(ends.min(1)))
    print("   per CTA: last warp to finish  ", rng(ends.max(1)))
    print("   epilogue (last warp -> CTA end)", rng(s[:, 20] - t0 - ends.max(1)))
    print("   CTA end                       ", rng(s[:, 20] - t0))
    by_smsp = [np.median(ends[:, [w for w in range(nw) if w % 4 == q]]) for q in range(4)]
    print("   median warp end by scheduler (warp %% 4): " + " ".join("%.1f" % v for v in by_smsp))
    print("   median warp end by warp index: " + " ".join("%.0f" % np.median(ends[:, w]) for w in range(nw)), flush=True)
    np.savetxt(os.path.join(os.environ.get("HPV_STAMP_OUT", "."), "bwd_stamps_%s.csv" % name), s - t0 * (s > 0), fmt="%.2f", delimiter=",")
    eng.close()
