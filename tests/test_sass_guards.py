"""CPU: guards on the generated SASS of the two dominant kernels (needs cuobjdump and the in-tree build).

The layer products read their weights from constant memory through the uniform datapath (`LDCU.64` into uniform
registers, `FFMA2` with a UR operand).  ptxas silently falls back to vector-indexed `LDC.64` + register operands
-- about 1.5x slower -- as soon as it cannot prove the loop control around the products warp-uniform (a per-warp trip
count, a thread-dependent branch ahead of the sweep: DESIGN.md section 4).  These tests catch that regression without
a GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "build", "hpv")


def _sass(obj, mangled_fragment):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    path = os.path.join(OBJ, obj)
    if not os.path.exists(exe) or not os.path.exists(path):
        pytest.skip("cuobjdump or %s not available (run __graft_entry__.build())" % obj)
    out = subprocess.run([exe, "-sass", path], capture_output=True, text=True, check=True).stdout
    blocks = out.split("Function : ")
    sel = [b for b in blocks if b.startswith(mangled_fragment)]
    assert len(sel) == 1, "kernel %s not found in %s" % (mangled_fragment, obj)
    return sel[0]


def _weight_loads(sass):
    uniform = len(re.findall(r"LDCU\.(?:64|128) UR\d+, c\[0x3\]", sass))
    vector = len(re.findall(r"LDC\.64 R\d+, c\[0x3\]\[R", sass))
    ur_ffma2 = len([l for l in sass.splitlines() if "FFMA2" in l and "UR" in l])
    return uniform, vector, ur_ffma2


def test_reverse_sweep_keeps_the_uniform_datapath():
    # hpv_mlpbwd_kernel<2,1,0,20,tanh>: the directional reverse sweep of the headline configuration
    s = _sass("hpv_k_h20_bwd.o", "_Z17hpv_mlpbwd_kernelILi2ELi1ELi0ELi20ELi1EEv10HpvBwdArgs")
    uniform, vector, ur_ffma2 = _weight_loads(s)
    assert uniform >= 100, "weight loads of the layer products are no longer LDCU (%d uniform, %d vector)" % (uniform, vector)
    assert vector <= 40        # the once-per-tile loads (bias, first and output layer) may be vector-indexed
    assert ur_ffma2 >= 150
    m = re.search(r"REG:(\d+)", subprocess.run([shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump", "-res-usage",
                                                os.path.join(OBJ, "hpv_k_h20_bwd.o")], capture_output=True, text=True).stdout.split(
        "hpv_mlpbwd_kernelILi2ELi1ELi0ELi20ELi1E")[1])
    assert m and int(m.group(1)) <= 128, "the two-channel reverse sweep must fit 128 registers (15-16 warps per SM)"
    assert "BAR.SYNC" in s and s.count("BAR.SYNC") <= 8, "no CTA barrier inside the sweep (prologue/epilogue only)"


def test_forward_kernel_keeps_the_uniform_datapath():
    s = _sass("hpv_k_h20_fwd.o", "_Z17hpv_varfwd_kernelILi2ELi1ELi1ELi20ELi1EEv10HpvVarArgs")
    uniform, vector, ur_ffma2 = _weight_loads(s)
    assert uniform >= 40 and ur_ffma2 >= 100, (uniform, vector, ur_ffma2)
    assert vector <= 30


def test_tensor_core_kernels_use_tcgen05_tmem_and_the_bulk_copy_engine():
    """BASELINE.json north_star: MLP layer products on tensor-core tiles, tables staged through TMA.  The headline
    instances of the tensor-core forward kernel and reverse sweep must contain the Blackwell instructions
    (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk -> UBLKCP) and issue their MMAs densely
    (operands in uniform registers: few R2UR transfers per MMA)."""
    fwd = _sass("hpv_k_h20_fwdtc.o", "_Z20hpv_varfwd_tc_kernelILi2ELi1ELi1ELi20ELi1EEv10HpvVarArgs")
    n_mma = len(re.findall(r"UTC\w*MMA", fwd))
    assert n_mma >= 54, n_mma                           # 3 channels x 9 MMAs, for the two TMEM base addresses
    assert "LDTM" in fwd and "STTM" in fwd and "UBLKCP" in fwd
    assert fwd.count("R2UR") <= 4 * n_mma
    bwd = _sass("hpv_k_h20_bwdtc.o", "_Z21hpv_mlpbwd_tcw_kernelILi2ELi1ELi0ELi20ELi1EEv10HpvBwdArgs")
    n_mma_b = len(re.findall(r"UTC\w*MMA", bwd))
    assert n_mma_b >= 2 * (18 + 32), n_mma_b            # products (2 channels x 9) + weight gradients (16 K steps x 2), two bases
    assert "LDTM" in bwd and "STTM" in bwd
