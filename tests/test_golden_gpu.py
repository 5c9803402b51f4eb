"""GPU: the CUDA path through the C ABI against the committed golden vectors (no oracle involved)."""
import pytest

import golden_util as G
from helpers import assert_result_parity, assert_same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,every", G.cases())
def test_cuda_reproduces_golden(pcf, name, every):
    fx = G.load(name)
    fus = pcf.Fusion(fx.box, fx.res, fx.clip[0], fx.clip[1])
    G.replay(fus, fx, every)
    assert fus.dims == fx.dims
    assert_result_parity(fus.extract(), fx.result[every], f"{name}/{every} result.")
    assert_same(fus.state(), fx.state[every], G.STATE_F, f"{name}/{every} state.")
    fus.close()
