"""CPU: the product's host-side quadrature module against tables produced by the reference's own
GaussJacobiQuadRule_V3 / VPINN.Test_fcn / dTest_fcn (tests/golden/tables.npz) and known-answer identities."""
import os

import numpy as np
import pytest

from hpv_b200 import GaussJacobiQuadRule_V3 as GJ
from tests import _cases as C

TAB = dict(np.load(os.path.join(C.GOLDEN, "tables.npz")))


@pytest.mark.parametrize("Q", [5, 10, 50, 80])
def test_gll_rule_matches_reference_module(Q):
    x, w = GJ.GaussLobattoJacobiWeights(Q, 0, 0)
    assert np.allclose(x, TAB["gll_x_%d" % Q], rtol=0, atol=1e-14)
    assert np.allclose(w, TAB["gll_w_%d" % Q], rtol=1e-12, atol=1e-16)
    assert x[0] == -1.0 and x[-1] == 1.0 and abs(w.sum() - 2.0) < 1e-13
    # exactness: a Q-point Lobatto rule integrates polynomials up to degree 2Q-3
    for d in (0, 2, 2 * Q - 4):
        assert abs(np.sum(w * x ** d) - 2.0 / (d + 1)) < 1e-12


def test_test_function_tables_match_reference_class():
    x = TAB["gll_x_80"]
    T = GJ.Test_fcn(60, x)
    D1, D2 = GJ.dTest_fcn(60, x)
    assert np.abs(T - TAB["T_60_80"]).max() <= 1e-12 * np.abs(TAB["T_60_80"]).max()
    assert np.abs(D1 - TAB["D1_60_80"]).max() <= 1e-12 * np.abs(TAB["D1_60_80"]).max()
    assert np.abs(D2 - TAB["D2_60_80"]).max() <= 1e-11 * np.abs(TAB["D2_60_80"]).max()
    b1, b2 = GJ.dTest_fcn(60, np.array([-1.0, 1.0]))
    assert np.allclose(b1, TAB["D1_bound_60"], rtol=1e-12) and np.allclose(b2, TAB["D2_bound_60"], rtol=1e-12)
    assert np.abs(T[:, [0, -1]]).max() < 1e-13                     # phi_n(+-1) = 0


def test_djacobi_identity():
    x = np.linspace(-0.9, 0.9, 7)
    for n in (1, 2, 5, 12):
        lhs = GJ.DJacobi(n + 1, 0, 0, x, 1) - GJ.DJacobi(n - 1, 0, 0, x, 1)
        assert np.allclose(lhs, GJ.dTest_fcn(n, x)[0][n - 1], rtol=1e-12, atol=1e-12)
    assert np.all(GJ.DJacobi(0, 0, 0, x, 1) == 0)
    xg, wg = GJ.GaussJacobiWeights(6, 0, 0)
    assert abs(np.sum(wg * xg ** 10) - 2.0 / 11) < 1e-13
