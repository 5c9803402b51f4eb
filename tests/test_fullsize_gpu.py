"""BASELINE.json's configurations at (or near) full size on the GPU.

C1 (20 x 640x480, 1 mm, 0.5 m box) is compared bit for bit with the CPU oracle in full.  The larger grids
(1 m box @ 1 mm and 0.5 m box @ 0.5 mm = 1000^3 cells, where the oracle's dense CPU grid would need > 16 GB) are
covered through size-independent properties: strict x-major order of the extraction, conservation of points
(sum of per-voxel buffer lengths == points that passed clip + crop), idempotent extraction, clear() -> empty, and
equality of a frame-sharded run (ranks emulated as contexts) with the single-context run."""
import importlib

import numpy as np
import pytest

from helpers import RESULT_FIELDS, STATE_FIELDS, assert_result_parity, assert_same

pytestmark = pytest.mark.gpu


def _synth(pcf):
    return importlib.import_module(pcf.__name__ + ".synth")


def _push_all(fus, scene, frames, first=0, batch=64):
    import torch
    pts = np.stack([scene.frame(i)[0] for i in frames])
    poses = np.stack([scene.frame(i)[1] for i in frames])
    dev = torch.from_numpy(pts).cuda()
    for b in range(0, len(frames), batch):
        k = min(batch, len(frames) - b)
        fus.push_frames_device(dev[b], k, scene.points_per_frame, 4, poses[b:b + k], first + b)
    fus.sync()
    return pts, poses


def _properties(fus, kept, canonical=True):
    res = fus.extract()
    h = res.hash
    assert len(h) > 1000
    assert np.all(h[1:] > h[:-1]), "extraction is not in strict x-major order"
    st = fus.state()
    assert np.all(st.hash[1:] > st.hash[:-1])
    if canonical:
        assert int(st.buffer_len.sum()) == kept, "points were lost or duplicated between the log and the voxel buffers"
    assert int(st.normal_found.sum()) >= len(h)          # pad cells can hold a normal but are never exported
    assert np.all(res.count >= 0) and np.all(np.isfinite(res.normal))
    n = np.linalg.norm(res.normal.astype(np.float64), axis=1)
    assert np.all(np.abs(n - 1.0) < 1e-3)
    assert_same(fus.extract(), res, RESULT_FIELDS, "second extraction: ")
    return res, st


def test_c1_replay20_bit_exact_vs_oracle(pcf, oracle):
    scene = _synth(pcf).sphere_turntable(20)
    g = scene.grid
    fus = pcf.Fusion(g.box, g.res, log_capacity_hint=20 * scene.points_per_frame)
    og = oracle.OracleGrid(g.box, g.res)
    pts, poses = _push_all(fus, scene, range(20))
    kept = 0
    for i in range(20):
        kept += og.add_frame(pts[i], poses[i])
    assert fus.count_kept() == kept
    fus.update(); og.update()
    assert fus.dims == og.dims == (499, 499, 499)
    want = og.download()
    assert len(want) > 100000
    assert_result_parity(fus.extract(), want, "C1 result.")
    assert_same(fus.state(), og.state(), STATE_FIELDS, "C1 state.")
    fus.clear()
    assert len(fus.extract()) == 0 and fus.count_kept() == 0
    fus.close()


def test_c3_one_metre_box_properties_and_sharding(pcf):
    """1 m box @ 1 mm = 1000^3 cells (4 GB grid): 24 frames of the plate sweep; sharded x3 == single context."""
    sh = importlib.import_module(pcf.__name__ + ".sharded")
    scene = _synth(pcf).plate_sweep(1000)
    g = scene.grid
    frames = list(range(0, 24))
    one = pcf.Fusion(g.box, g.res)
    assert one.dims == (999, 999, 999)
    _push_all(one, scene, frames)
    kept = one.count_kept()
    assert kept > 0.5 * len(frames) * scene.points_per_frame
    one.update()
    res, _ = _properties(one, kept)
    ranks = [pcf.Fusion(g.box, g.res) for _ in range(3)]
    for r, f in enumerate(ranks):
        lo, hi = sh.frame_block(len(frames), r, 3)
        _push_all(f, scene, frames[lo:hi], first=lo)
    got = sh.merge_and_extract_local_v2(ranks)
    assert_same(got, res, RESULT_FIELDS, "C3 sharded x3 (exchange v2): ")
    for f in ranks + [one]:
        f.close()


def test_c4_hires_half_millimetre_interleaved_properties(pcf):
    """1920x1080 clouds, 0.5 mm voxels (999^3), update after every 2 frames: the incremental dependants path at size."""
    scene = _synth(pcf).hires_sphere(6)
    g = scene.grid
    fus = pcf.Fusion(g.box, g.res)
    assert fus.dims == (999, 999, 999)
    import torch
    for i in range(6):
        pts, T = scene.frame(i)
        fus.push_frames_device(torch.from_numpy(pts).cuda(), 1, scene.points_per_frame, 4, T[None], i)
        if i % 2 == 1:
            fus.update()
    kept = fus.count_kept()
    fus.update()
    res, st = _properties(fus, kept, canonical=False)
    assert int(st.buffer_len.sum()) < kept      # points landing in voxels that already have a normal are not buffered (OG.hpp:210-216)
    assert int(res.count.sum()) > 0
    fus.close()


def test_c5_sheets_world_points_bit_exact_vs_oracle(pcf, oracle):
    """Extraction stress shape at oracle-friendly size: stacked wavy sheets inserted through pcf_add_points
    (OccupancyGrid::addPoints semantics: world-frame points + explicit viewpoint)."""
    g, sheets = _synth(pcf).wavy_sheets_world(n_sheets=3, n_side=160)
    box = (-0.1, 0.1, -0.1, 0.1, -0.5, 0.5)     # a thin column keeps the oracle's dense grid small
    fus = pcf.Fusion(box, g.res)
    og = oracle.OracleGrid(box, g.res)
    for i, (pts, vp) in enumerate(sheets):
        fus.add_points(pts, vp, i)
        og.add_points_world(pts, vp)
    fus.update(); og.update()
    want = og.download()
    assert len(want) > 20000
    assert_result_parity(fus.extract(), want, "C5 result.")
    assert_same(fus.state(), og.state(), STATE_FIELDS, "C5 state.")
    fus.close()


def test_c4_resolution_interleaved_bit_exact_on_a_small_box(pcf, oracle):
    """C4's ingredients -- 1920x1080 clouds, 0.5 mm voxels, an update between frames -- on a box small enough for the oracle's
    dense CPU grid (a 4 cm sphere in a 12 cm box: 240^3 cells): bit-exact state and extraction."""
    synth = _synth(pcf)
    scene = synth.sphere_turntable(6, 1920, 1080, 0.0005, fx=1800.0, radius=0.04, standoff=0.36, box_half=0.06)
    g = scene.grid
    fus, og = pcf.Fusion(g.box, g.res), oracle.OracleGrid(g.box, g.res)
    assert fus.dims == og.dims == (239, 239, 239)
    kept = 0
    for i in range(scene.n_frames):
        pts, T = scene.frame(i)
        fus.push_frame(pts, T, i)
        kept += og.add_frame(pts, T)
        if i in (1, 3):
            fus.update(); og.update()
    assert fus.count_kept() == kept and kept > 300000
    fus.update(); og.update()
    want = og.download()
    assert len(want) > 50000
    assert_result_parity(fus.extract(), want, "C4-small result.")
    assert_same(fus.state(), og.state(), STATE_FIELDS, "C4-small state.")
    fus.close()
