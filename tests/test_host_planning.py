"""CPU: host-side planning shared by the C ABI and the kernels (hp-vpinns_b200/csrc/hpv_host_prep.h, hpv_varbwd.cuh),
through the emulation library: the forward kernel's work partition and the compact gradient layout of the reverse
sweep."""
import ctypes

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from tests import _emu as E


def _partition(n_el, pts, tile, max_ctas, cta_pts):
    L = E.lib()
    tb = np.zeros(max_ctas + 2, dtype=np.int32)
    first = np.zeros(n_el, dtype=np.int32); nparts = np.zeros(n_el, dtype=np.int32); off = np.zeros(n_el, dtype=np.int32)
    tpe = ctypes.c_int(0); total = ctypes.c_int(0)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    n = L.hpv_emu_partition(n_el, pts, tile, max_ctas, cta_pts, ctypes.byref(tpe), ctypes.byref(total), ip(tb), ip(first), ip(nparts), ip(off))
    return n, tpe.value, total.value, tb[:n + 1], first, nparts, off


@settings(max_examples=200, deadline=None)
@given(n_el=st.integers(1, 40), pts=st.integers(1, 7000), tile=st.sampled_from([32, 256]), max_ctas=st.integers(1, 300),
       full=st.booleans())
def test_forward_partition_invariants(n_el, pts, tile, max_ctas, full):
    n, tpe, total, tb, first, nparts, off = _partition(n_el, pts, tile, max_ctas, 256 if full else 0)
    ntiles = n_el * tpe
    assert tpe == -(-pts // tile) and 1 <= n <= max_ctas
    assert tb[0] == 0 and tb[-1] == ntiles and np.all(np.diff(tb) >= 0)                 # contiguous cover
    sizes = np.diff(tb)
    assert sizes.max() - sizes.min() <= 1                                                # balanced to one tile
    if ntiles >= n:
        assert sizes.min() >= 1
    # the CTAs that touch an element are exactly first .. first + nparts - 1, and the partial-U slots are packed
    acc = 0
    for e in range(n_el):
        t0, t1 = e * tpe, (e + 1) * tpe - 1
        touching = [c for c in range(n) if tb[c] <= t1 and tb[c + 1] > t0]
        assert touching == list(range(first[e], first[e] + nparts[e]))
        assert off[e] == acc
        acc += nparts[e]
    assert total == acc


def _cost(tb, tpe, cross, first_wave, wave2):
    out = []
    for c in range(len(tb) - 1):
        alpha = wave2 if (first_wave > 0 and c >= first_wave) else 1.0
        inside = sum(1 for t in range(tb[c] + 1, tb[c + 1]) if t % tpe == 0)
        out.append(alpha * (tb[c + 1] - tb[c]) + cross * inside)
    return np.array(out)


@settings(max_examples=100, deadline=None)
@given(n_el=st.integers(1, 70), pts=st.integers(100, 7000), max_ctas=st.integers(2, 300), cross=st.floats(0.0, 3.0),
       wave2=st.floats(1.0, 1.3))
def test_weighted_forward_partition_covers_and_does_not_increase_the_modelled_maximum(n_el, pts, max_ctas, cross, wave2):
    L = E.lib()
    tile = 128
    tb = np.zeros(max_ctas + 2, dtype=np.int32); tpe = ctypes.c_int(0)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    L.hpv_emu_partition_weighted.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    first_wave = max_ctas // 2
    n = L.hpv_emu_partition_weighted(n_el, pts, tile, max_ctas, cross, first_wave, wave2, ctypes.byref(tpe), ip(tb))
    tb = tb[:n + 1]; tpe = tpe.value
    ntiles = n_el * tpe
    assert tb[0] == 0 and tb[-1] == ntiles and np.all(np.diff(tb) >= (1 if ntiles >= n else 0))      # contiguous cover, nobody idle
    # against equal tile counts the largest modelled cost does not grow
    uniform = np.array([(ntiles * c) // n for c in range(n + 1)])
    # (the assignment rounds every CTA to the cost nearest the running target: allow one rounding step)
    assert _cost(tb, tpe, cross, first_wave, wave2).max() <= _cost(uniform, tpe, cross, first_wave, wave2).max() + wave2 + cross + 1e-6


def test_weighted_partition_of_the_headline_case_relieves_the_ctas_that_straddle_elements():
    L = E.lib()
    tb = np.zeros(298, dtype=np.int32); tpe = ctypes.c_int(0)
    ip = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))
    L.hpv_emu_partition_weighted.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int, ctypes.c_double,
                                                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    n = L.hpv_emu_partition_weighted(64, 6400, 128, 296, 2.4, 148, 1.16, ctypes.byref(tpe), ip(tb))     # C3 on 148 SMs x 2
    assert n == 296 and tpe.value == 50
    tb = tb[:n + 1]
    cost = _cost(tb, 50, 2.4, 148, 1.16)
    uniform = np.array([(3200 * c) // n for c in range(n + 1)])
    assert cost.max() < _cost(uniform, 50, 2.4, 148, 1.16).max() - 1.0          # 15.2 -> about 12.6 tile times
    assert cost.min() > cost.max() - 3.5                                      # nobody idles at the end of the list
    sizes = np.diff(tb)
    assert sizes[:148].mean() > sizes[148:].mean()                            # the second CTA of an SM gets less


@pytest.mark.parametrize("dim", [1, 2])
@pytest.mark.parametrize("hp", [8, 20, 32])
@pytest.mark.parametrize("nhid", [1, 2, 3, 8])
def test_compact_gradient_layout_is_a_bijection_onto_the_primary_parameters(dim, hp, nhid):
    """Every padded slot except the transposed copies of the hidden matrices maps to exactly one compact slot."""
    L = E.lib()
    npad = L.hpv_emu_theta_pad_n(dim, hp, nhid)
    ncompact = L.hpv_emu_gw_n(dim, hp, nhid)
    transposed = set()
    for l in range(1, nhid):
        o = L.hpv_emu_off_wt(dim, hp, l)
        transposed |= set(range(o, o + hp * hp))
    seen = []
    for ip in range(npad):
        ic = L.hpv_emu_gw_of_padded(dim, hp, nhid, ip)
        if ip in transposed:
            assert ic == -1
        else:
            assert 0 <= ic < ncompact
            seen.append(ic)
    assert sorted(seen) == list(range(ncompact))
    # hidden matrix l sits at the same relative position in both layouts
    for l in range(1, nhid):
        assert L.hpv_emu_gw_of_padded(dim, hp, nhid, L.hpv_emu_off_wl(dim, hp, l)) == (dim + 1) * hp + (l - 1) * (hp * hp + hp)


def _tc_plan(dim, hp, nch, nch1, nhid, Q, rows, n_terms, ltab=0, rtab=1):
    L = E.lib()
    out = np.zeros(19, dtype=np.int32)
    need = L.hpv_emu_tc_plan(dim, hp, nch, nch1, nhid, Q, rows, n_terms, ltab, rtab, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return need, out


def test_tensor_core_plans_of_the_headline_configuration():
    """C3/C4: [2,20,20,20,1], value + two tangents forward, directional reverse sweep.  The tensor-core kernels need two
    CTAs per SM: <= 256 TMEM columns and <= 113 KB of shared memory each; the weight tiles and mbarriers must sit
    at the alignments the hardware asks for."""
    need, o = _tc_plan(2, 20, 3, 3, 3, 80, 80, 2)
    assert need == 3 * 32 + 2 * 3 * 24 == 240                     # D + A_hi + A_lo
    assert o[0] * 4 <= 113 * 1024                                 # forward: two CTAs per SM
    assert o[1] % 32 == 0 and o[2] % 2 == 0                       # B tiles 128-byte aligned, mbarriers 8-byte aligned
    assert o[5] % 4 == 0                                          # tables: 16-byte aligned destination of the bulk copy
    assert o[1] + 2 * 2 * 24 * 32 <= o[2]                         # two layers of (hi, lo) weight tiles before the barriers
    need_b, ob = _tc_plan(2, 20, 2, 2, 3, 80, 80, 2)              # directional reverse sweep: 2 channels
    assert need_b == 160
    assert ob[8] * 4 <= 113 * 1024 and ob[9] % 32 == 0 and ob[10] % 2 == 0
    assert ob[14] * 4 <= 113 * 1024 and ob[15] % 32 == 0 and ob[16] % 2 == 0
    L = E.lib()
    assert L.hpv_emu_tcw_tmem_need(2, 20, 3) == 160 + 2 * 32 <= 256
    assert L.hpv_emu_tcw_supported(2, 20, 3) == 1 and L.hpv_emu_tcw_supported(2, 32, 3) == 0 and L.hpv_emu_tcw_supported(2, 20, 1) == 0


@settings(max_examples=100, deadline=None)
@given(hp=st.sampled_from([8, 20, 32]), nch=st.integers(1, 5), nhid=st.integers(1, 8), Q=st.integers(2, 128), n_terms=st.integers(1, 2),
       dim=st.sampled_from([1, 2]))
def test_tensor_core_plan_invariants(hp, nch, nhid, Q, n_terms, dim):
    rows = Q if dim == 2 else 1
    need, o = _tc_plan(dim, hp, nch, min(nch, 3), nhid, Q, rows, n_terms)
    kp = ((hp + 1 + 7) // 8) * 8
    assert need == nch * 32 + 2 * nch * kp
    for base in (0, 8, 14):                                       # the three kernels' plans
        total, B, bar = o[base], o[base + 1], o[base + 2]
        assert 0 < B <= bar < total and B % 32 == 0 and bar % 2 == 0
    # regions of the forward plan do not overlap: G < tables < P < th < part < B
    assert o[7] < o[5] < o[6] <= o[3] < o[4] < o[1]
