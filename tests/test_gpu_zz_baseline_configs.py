"""GPU parity (-m gpu) at the full sizes of the BASELINE.json configurations that are parity cases: C1, C2 (1-D
Poisson) and C5 (AdvDiff identification), through the C ABI against the float64 oracle.  (C3 and a C4 shard are in
test_gpu_fullsize.py.)  Same tolerances as test_gpu_parity.py."""
import numpy as np
import pytest

from tests import _baseline_cases as B
from tests import _gpu as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1", "c2", "c5"])
def test_baseline_config_at_full_size(name):
    inp, l_ref, g_ref, r_ref, ge_ref = B.build(name)
    eng = G.make_engine(inp)
    loss, res = eng.varloss_forward()
    g, ge = eng.varloss_backward()
    assert loss == pytest.approx(l_ref, rel=1e-5)
    assert np.abs(res - r_ref.reshape(res.shape)).max() <= 2e-5 * np.abs(r_ref).max()
    assert np.abs(g - g_ref).max() <= 1e-4 * np.abs(g_ref).max()
    if ge_ref is not None:
        assert ge == pytest.approx(ge_ref, rel=1e-4)
    loss2, res2 = eng.varloss_forward()
    g2, _ = eng.varloss_backward()
    assert loss2 == loss and np.array_equal(res, res2) and np.array_equal(g, g2)          # deterministic
    eng.close()
