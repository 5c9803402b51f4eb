"""CPU: the oracle restatement against the committed golden vectors (tests/golden, made by tools/make_golden.py from
the reference's own OccupancyGrid.hpp compiled against the shim headers), and -- where the prebuilt oracle/_ref
libraries are present -- the reference build itself against the same vectors."""
import numpy as np
import pytest

import golden_util as G
from helpers import assert_same


@pytest.mark.parametrize("name,every", G.cases())
def test_oracle_reproduces_golden(oracle, name, every):
    fx = G.load(name)
    og = oracle.OracleGrid(fx.box, fx.res, fx.clip[0], fx.clip[1])
    G.replay(og, fx, every)
    assert og.dims == fx.dims
    assert_same(og.download(), fx.result[every], G.RES_F, f"{name}/{every} result.")
    assert_same(og.state(), fx.state[every], G.STATE_F, f"{name}/{every} state.")


@pytest.mark.parametrize("name,every", G.cases())
def test_reference_build_reproduces_golden(oracle, name, every, capfd):
    if not oracle.available("ref_ordered"):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    fx = G.load(name)
    og = oracle.OracleGrid(fx.box, fx.res, fx.clip[0], fx.clip[1], kind="ref_ordered")
    G.replay(og, fx, every)
    assert_same(og.download(), fx.result[every], G.RES_F, f"{name}/{every} result.")
    got = og.state()
    for f in ["hash", "buffer_len", "normal_found", "count"]:
        assert np.array_equal(getattr(got, f), getattr(fx.state[every], f))


def test_golden_is_nontrivial():
    for name in G.fixtures():
        fx = G.load(name)
        r = fx.result[fx.schedules[0]]
        assert len(r.hash) > 5000 and r.count.sum() > 1000
        assert np.all(np.diff(r.hash.astype(np.int64)) > 0)          # x-major extraction order == ascending hash
        n = np.linalg.norm(r.normal.astype(np.float64), axis=1)
        assert np.all(np.abs(n[np.isfinite(n)] - 1) < 1e-5)
