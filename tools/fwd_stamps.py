"""Timing experiment (development aid): phase timeline of the tensor-core forward kernel, from a library built with
HPV_NVCC_EXTRA=-DHPV_EXP_STAMPS (HPV_LIB points to it).  Thread 0 of every CTA records %globaltimer at the phase
boundaries and the time spent per phase; this prints, over the CTAs of the last launch, the spread of each."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

NST = 12
lib = ctypes.CDLL(os.environ["HPV_LIB"])
lib.hpv_exp_read_stamps.argtypes = [ctypes.c_void_p, ctypes.c_int]
for ne in (2, 8, 11, 16):
    bench.WORKLOADS["x"] = dict(bench.WORKLOADS["c3"], desc="x", ne=ne)
    wl = bench.build_workload("x")
    eng = bench.make_engine(wl, 0)
    t = eng.time_kernel(0, 20)
    info = eng.kernel_info()
    n = info["fwd_grid"]
    buf = np.zeros((n, NST), dtype=np.uint64)
    rc = lib.hpv_exp_read_stamps(buf.ctypes.data, n)
    assert rc == 0, rc
    s = buf.astype(np.float64) * 1e-3          # us
    t0 = s[:, 0].min()
    def rng(v): return "min %6.1f med %6.1f max %6.1f" % (v.min(), np.median(v), v.max())
    print("n_el %4d  fwd %.1f us  grid %d" % (ne * ne, t, n))
    print("   CTA start after first CTA   ", rng(s[:, 0] - t0))
    print("   set-up (entry -> ready)     ", rng(s[:, 1] - s[:, 0]))
    print("   wait for the tables         ", rng(s[:, 2] - s[:, 1]))
    print("   tile loop total             ", rng(s[:, 3] - s[:, 2]))
    print("     MLP phase                 ", rng(s[:, 5]))
    print("     contraction over x        ", rng(s[:, 6]))
    print("     contraction over y        ", rng(s[:, 7]))
    print("     element publish / finish  ", rng(s[:, 8]))
    print("   exit (dealloc)              ", rng(s[:, 4] - s[:, 3]))
    print("   CTA end after first start   ", rng(s[:, 4] - t0), flush=True)
    if ne == 8:
        np.savetxt(os.path.join(os.environ.get('HPV_STAMP_OUT', '.'), 'fwd_stamps_c3_per_cta.csv'), s - t0 * (np.arange(NST) < 5), fmt='%.2f', delimiter=',')
    eng.close()
