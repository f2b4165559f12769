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
