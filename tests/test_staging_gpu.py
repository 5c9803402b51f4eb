"""Host staging pool (pcf_submit_* / pcf_stage_frame, csrc/pcf_stager.hpp) and the PointCloud2 layouts.

The clip-and-pack staging replaces the reference's addPoints() thread (node.cpp:218-263): it may only drop points the
depth clip drops anyway and must keep point order, so every grid built through it has to be bit-identical to the one the
unstaged float4 path builds (which tests/test_parity_gpu.py anchors to the oracle) -- and to the oracle directly."""
import os

import numpy as np
import pytest

from helpers import RESULT_FIELDS, STATE_FIELDS, assert_result_parity, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(pcf):
    import importlib
    return importlib.import_module(pcf.__name__ + ".synth").small_sphere(6)


@pytest.mark.parametrize("threads,raw_lanes,pinned", [(1, -1, False), (5, -1, False), (1, 2, True), (3, 1, True), (2, 0, False)])
@pytest.mark.parametrize("update_every", [None, 2])
def test_submit_frame_bit_identical_to_push_frame_and_oracle(pcf, oracle, small, threads, raw_lanes, pinned, update_every):
    """raw_lanes: -1 / 0 none; 2 / 1 with pinned clouds: some clouds are uploaded unstaged by the raw lanes while the packers are
    busy (which ones depends on timing -- the grid must not); pageable clouds are always packed."""
    import torch
    g = small.grid
    a = pcf.Fusion(g.box, g.res, stage_threads=threads, stage_raw_lanes=raw_lanes)
    og = oracle.OracleGrid(g.box, g.res)
    keep = []
    for rep in range(3):                       # 18 frames: more clouds than pinned slots, the ring wraps
        for i in range(small.n_frames):
            pts, T = small.frame(i)
            k = rep * small.n_frames + i
            pose = np.ascontiguousarray(T, np.float64).reshape(16)
            if pinned:
                pts_t = torch.from_numpy(pts).pin_memory()
                pts = pts_t.numpy()
                keep.append(pts_t)
            keep.append((pts, pose))           # submitted clouds are read asynchronously
            assert a.submit_frame(pts, pose, k) == 0
            og.add_frame(pts, T)
            if update_every and (k + 1) % update_every == 0:
                a.update(); og.update()
    a.update(); og.update()
    st = a.stats()
    assert st["frames_pushed"] == 18 and st["points_offered"] == 18 * small.points_per_frame
    if raw_lanes < 0 or not pinned:
        assert st["h2d_bytes"] < 18 * small.points_per_frame * 12  # only the clipped cloud crossed PCIe
    assert_result_parity(a.extract(), og.download(), "submit_frame result.")
    assert_same(a.state(), og.state(), STATE_FIELDS, "submit_frame state.")
    a.close()


def test_stage_frame_then_push_equals_plain_push(pcf, small):
    g = small.grid
    a, b = pcf.Fusion(g.box, g.res), pcf.Fusion(g.box, g.res)
    for i in range(small.n_frames):
        pts, T = small.frame(i)
        if i == 1:
            pts = pts.copy(); pts[::7, 2] = 0.7; pts[3::11, 2] = 0.1; pts[5::13, 2] = 0.28     # outside / on the clip limits
        a.push_frame(pts, T, i)
        staged, m = b.stage_frame(pts)
        assert m % 4 == 0 and m <= len(pts) + 3
        z = pts[:, 2].astype(np.float64)             # node.cpp:251 compares the float promoted to double
        inside = (z > 0.28) & (z < 0.6)
        assert m - int(inside.sum()) in (0, 1, 2, 3)
        assert np.array_equal(staged[:int(inside.sum())].view(np.uint32), pts[inside][:, :3].view(np.uint32))   # order kept
        b.push_frame(np.ascontiguousarray(staged), T, i)
    a.update(); b.update()
    assert a.count_kept() == b.count_kept()
    assert_same(b.extract(), a.extract(), RESULT_FIELDS, "stage_frame: ")
    assert_same(b.state(), a.state(), STATE_FIELDS, "stage_frame state: ")
    a.close(); b.close()


@pytest.mark.parametrize("x_offset,pad,submit", [(0, 0, False), (8, 0, False), (4, 24, False), (0, 12, True), (8, 20, True)])
def test_pointcloud2_offsets_and_padded_rows(pcf, small, x_offset, pad, submit):
    """x/y/z at a non-zero field offset and rows with trailing padding (node.cpp:182-216 reads fields[i].offset and
    row_step): both the strided upload (pcf_push_pointcloud2 -- uploads exactly the message bytes, the message buffer ends
    right after the last point) and the staged route give the grid of the float4 path."""
    g = small.grid
    a, b = pcf.Fusion(g.box, g.res), pcf.Fusion(g.box, g.res)
    W, H = small.width, small.height
    step = 32
    row = W * step + pad
    msgs = []
    for i in range(3):
        pts, T = small.frame(i)
        a.push_frame(pts, T, i)
        n_bytes = (H - 1) * row + W * step                           # no padding after the last row: an over-read would leave the buffer
        msg = np.full(n_bytes, 0x7f, np.uint8)
        for r in range(H):
            v = msg[r * row: r * row + W * step].reshape(W, step)
            v[:, x_offset:x_offset + 12] = pts[r * W:(r + 1) * W, :3].copy().view(np.uint8).reshape(W, 12)
        msgs.append(msg)
        off = (x_offset, x_offset + 4, x_offset + 8)
        if submit:
            assert b.submit_pointcloud2(msg, W, H, step, row, off, T, i) == 0
        else:
            assert b.push_pointcloud2(msg, W, H, step, row, off, T, i) == 0
    a.update(); b.update()
    assert_same(b.extract(), a.extract(), RESULT_FIELDS, "PointCloud2 layout: ")
    assert_same(b.state(), a.state(), STATE_FIELDS, "PointCloud2 layout state: ")
    a.close(); b.close()


def test_reset_closes_the_gate_and_drops_only_untaken_clouds(pcf, small):
    """node.cpp:351-359: reset sets start_ = false and clears the input deque; clouds a staging thread already took still
    integrate; later clouds are dropped until start.  The grid is left alone."""
    g = small.grid
    f = pcf.Fusion(g.box, g.res, stage_threads=1)
    clouds = [small.frame(i) for i in range(small.n_frames)]
    poses = [np.ascontiguousarray(T, np.float64).reshape(16) for _, T in clouds]
    for i in range(4):
        assert f.submit_frame(clouds[i][0], poses[i], i) == 0
    f.wait_staged(1)                                                 # the staging thread has taken (and handed over) the first cloud
    assert f.staged_count() >= 1
    f.reset()
    assert f.submit_frame(clouds[4][0], poses[4], 4) == 1            # PCF_DROPPED: the gate is closed
    assert f.push_frame(clouds[4][0], clouds[4][1], 4) == 1
    f.sync()
    st = f.stats()
    assert st["frames_pushed"] + st["staged_dropped"] == 4 and st["frames_pushed"] >= 1
    kept_before = f.count_kept()
    assert kept_before > 0                                          # the grid was kept
    f.start()
    nxt = 5
    assert f.submit_frame(clouds[5][0], poses[5], nxt) == 0
    f.update()
    assert f.count_kept() > kept_before and len(f.extract()) > 0
    f.close()


def test_scoring_variants_are_bit_identical(pcf, oracle, small):
    """k_score's work-balanced voxel order and its batched cylinder tests change scheduling only: every variant equals the oracle."""
    import importlib
    scene = importlib.import_module(pcf.__name__ + ".synth").sphere_turntable(8, 320, 240, 0.002)
    g = scene.grid
    og = oracle.OracleGrid(g.box, g.res)
    frames = [scene.frame(i) for i in range(scene.n_frames)]
    for pts, T in frames:
        og.add_frame(pts, T)
    og.update()
    want = og.download()
    assert len(want) > 20000
    try:
        for unr, bal, coop, slots in [(1, 0, 0, 16), (1, 1, 0, 16), (2, 1, 0, 16), (4, 1, 0, 16), (4, 0, 0, 16), (1, 1, 1, 16), (1, 0, 1, 16), (1, 1, 1, 8)]:
            os.environ["PCF_SCORE_UNR"], os.environ["PCF_SCORE_BALANCE"], os.environ["PCF_SCORE_COOP"] = str(unr), str(bal), str(coop)
            os.environ["PCF_COOP_SLOTS"] = str(slots)
            f = pcf.Fusion(g.box, g.res)
            for i, (pts, T) in enumerate(frames):
                f.push_frame(pts, T, i)
            f.update()
            assert_result_parity(f.extract(), want, f"k_score unroll={unr} balance={bal} coop={coop} slots={slots}: ")
            f.close()
    finally:
        for k in ("PCF_SCORE_UNR", "PCF_SCORE_BALANCE", "PCF_SCORE_COOP", "PCF_COOP_SLOTS"):
            os.environ.pop(k, None)


def test_shared_reciprocal_division_is_the_ieee_division(pcf, small):
    """k_score_coop divides six numbers by float(count) with ONE reciprocal (csrc/pcf_kernels.cuh div_shared): bit-identical to
    x / c (the compiler's div.rn.f32) for every integer divisor up to 2^17, random ones up to 2^24, and dividends covering
    every exponent, both zeros, denormals, infinities, NaNs and the magnitudes the fold really sees."""
    f = pcf.Fusion(small.grid.box, small.grid.res)
    rng = np.random.default_rng(11)
    n_div = 1 << 17
    for rep in range(6):
        c = np.arange(1, n_div + 1, dtype=np.float32)
        if rep >= 3:
            c = rng.integers(1, 1 << 24, n_div).astype(np.float32)
        if rep % 3 == 0:          # any bit pattern
            x = rng.integers(0, 1 << 32, n_div, dtype=np.uint64).astype(np.uint32).view(np.float32)
        elif rep % 3 == 1:        # the fold's magnitudes: coordinate differences and variance terms
            x = (rng.standard_normal(n_div) * 10.0 ** rng.uniform(-12, 0, n_div)).astype(np.float32)
        else:                     # quotients that land next to rounding boundaries: x = c * (k + 0.5 ulp-ish)
            k = rng.integers(1, 1 << 23, n_div).astype(np.float64)
            x = (c.astype(np.float64) * (k + rng.choice([0.5, 0.4999999, 0.5000001], n_div)) * 2.0 ** rng.integers(-60, 20, n_div)).astype(np.float32)
        special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3e38, 2.0 ** -80, 2.0 ** 80, 2.0 ** -81], np.float32)
        x[:len(special)] = special
        m, bad = f.kat_div(x, c)
        assert m == 0, f"rep {rep}: {m} quotients differ, e.g. x={x[bad]!r} c={c[bad]!r}"
    f.close()


def test_reserve_process_presizes_scratch_without_changing_results(pcf, small):
    """pcf_reserve_process: the first update / extract after it runs on pre-sized buffers (a live node's first process());
    results are the same bytes, and the first extraction is not slower than a warm one by the allocation cost."""
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    a, b = pcf.Fusion(g.box, g.res), pcf.Fusion(g.box, g.res)
    b.reserve_process(small.n_frames * small.points_per_frame, 200_000)
    for f in (a, b):
        for i, (pts, T) in enumerate(frames):
            f.push_frame(pts, T, i)
        f.update()
    ra, rb = a.extract(), b.extract()
    assert_same(rb, ra, RESULT_FIELDS, "reserve_process: ")
    assert_same(b.state(), a.state(), STATE_FIELDS, "reserve_process state: ")
    a.close(); b.close()
