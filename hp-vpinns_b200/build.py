"""Builds libhpv.so (the CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python hp-vpinns_b200/build.py [--force] [--jobs N]

One object per translation unit under build/ (git-ignored), compiled in parallel, re-used when neither the
source nor any header changed.  The shared library lands next to this file so that it travels to the GPU box
with the repository snapshot.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
BUILD = os.path.join(ROOT, "build", "hpv")
LIB = os.path.join(HERE, "libhpv.so")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


# extra nvcc flags for tuning experiments (e.g. HPV_NVCC_EXTRA="-DHPV_BWD_DIR_MIN_CTAS=2"); part of the object digest
NVCC_FLAGS += os.environ.get("HPV_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(paths):
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def build(force=False, jobs=None, verbose=True):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    kernel_headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC)
                      if f.endswith((".h", ".cuh")) and f != "hpv_host_prep.h"]
    api_headers = kernel_headers + [os.path.join(CSRC, "hpv_host_prep.h"), os.path.join(ROOT, "include", "hpv.h")]
    todo, objs = [], []
    for src in sources():
        path = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src[:-3] + ".o")
        stamp = obj + ".sha"
        dig = _digest((api_headers if src == "hpv_api.cu" else kernel_headers) + [path])
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        todo.append((path, obj, stamp, dig))

    def compile_one(item):
        path, obj, stamp, dig = item
        cmd = [nvcc] + NVCC_FLAGS + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (path, r.stdout, r.stderr))
        with open(stamp, "w") as f:
            f.write(dig)
        return os.path.basename(path)

    if todo:
        jobs = jobs or min(len(todo), os.cpu_count() or 4)
        if verbose:
            print("[hpv build] compiling %d translation units with %d jobs ..." % (len(todo), jobs), flush=True)
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            for name in ex.map(compile_one, todo):
                if verbose:
                    print("[hpv build]   %s" % name, flush=True)
    if todo or not os.path.exists(LIB) or force:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
        if verbose:
            print("[hpv build] linked %s" % LIB, flush=True)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--lib", default=None, help="output path of the shared library (default: next to this file)")
    a = ap.parse_args()
    if a.lib:
        LIB = os.path.abspath(a.lib)
        BUILD = BUILD + "_" + hashlib.sha256(LIB.encode()).hexdigest()[:8]
    build(force=a.force, jobs=a.jobs)
    sys.exit(0)
