"""CPU: the staging pool's clip-and-pack (csrc/pcf_pack.cpp; replaces node.cpp:182-216 decode + node.cpp:248-255 depth clip).
The scalar loop is the specification (keep iff clip_lo < z < clip_hi with the float thresholds equivalent to the reference's
double compares, NaN dropped, order kept, xyz packed); the AVX2 / AVX-512 loops must produce the same bytes for every
layout they accept."""
import numpy as np
import pytest

import pcfusion_b200 as pcf


def thresholds():
    lo = np.float32(0.28)
    if float(lo) > 0.28:
        lo = np.nextafter(lo, np.float32(-np.inf))      # largest float <= 0.28:   (double)z > 0.28  <=>  z > lo
    hi = np.float32(0.6)
    if float(hi) < 0.6:
        hi = np.nextafter(hi, np.float32(np.inf))       # smallest float >= 0.6:   (double)z < 0.6   <=>  z < hi
    return float(lo), float(hi)


def make_msg(rng, rows, cols, point_step, row_pad, x_offset):
    n = rows * cols
    xyz = rng.uniform(-1, 1, (n, 3)).astype(np.float32)
    xyz[:, 2] = rng.uniform(0.2, 0.7, n).astype(np.float32)
    xyz[rng.random(n) < 0.3] = np.nan                                   # background pixels
    edge = np.array([0.28, 0.6, np.nextafter(np.float32(0.28), np.float32(1)), np.nextafter(np.float32(0.6), np.float32(0)),
                     np.nextafter(np.float32(0.28), np.float32(0)), np.inf, -np.inf, 0.0], np.float32)
    idx = rng.integers(0, n, 64)
    xyz[idx, 2] = edge[rng.integers(0, len(edge), 64)]
    xyz[idx, 0] = 1.0
    xyz[idx, 1] = 2.0
    row_step = cols * point_step + row_pad
    msg = rng.integers(0, 255, (rows - 1) * row_step + cols * point_step, dtype=np.uint8)      # rgb / padding bytes: arbitrary
    for r in range(rows):
        v = msg[r * row_step: r * row_step + cols * point_step].reshape(cols, point_step)
        v[:, x_offset:x_offset + 12] = xyz[r * cols:(r + 1) * cols].copy().view(np.uint8).reshape(cols, 12)
    return msg, xyz, row_step


@pytest.mark.parametrize("point_step,x_offset,row_pad,rows,cols", [
    (16, 0, 0, 1, 1003), (16, 0, 0, 7, 333), (16, 4, 16, 5, 101), (32, 0, 0, 3, 257), (32, 8, 32, 4, 130), (32, 20, 0, 2, 99),
    (12, 0, 0, 1, 1001), (20, 0, 12, 6, 77), (16, 0, 0, 1, 3), (16, 0, 0, 1, 0)])
def test_clip_pack_all_implementations_agree_with_the_definition(point_step, x_offset, row_pad, rows, cols):
    rng = np.random.default_rng(point_step * 1000 + x_offset * 10 + rows)
    lo, hi = thresholds()
    msg, xyz, row_step = make_msg(rng, rows, max(cols, 1), point_step, row_pad, x_offset) if cols else (np.zeros(16, np.uint8), np.zeros((0, 3), np.float32), 16)
    z = xyz[:, 2].astype(np.float64)
    want = xyz[(z > 0.28) & (z < 0.6)]                                  # node.cpp:251 on doubles; NaN fails
    got = {}
    for isa in (0, 1, 2, -1, "2-unaligned"):     # 64-byte aligned and deliberately misaligned output buffers
        mis = isa == "2-unaligned"
        out, used = pcf.kat_clip_pack(msg, rows, cols, point_step, row_step, x_offset, lo, hi, 2 if mis else isa, misalign=mis)
        assert len(out) % 4 == 0 and 0 <= len(out) - len(want) < 4
        assert np.array_equal(out[:len(want)].view(np.uint32), want.view(np.uint32)), f"isa {isa} (ran {used})"
        assert np.all(np.isnan(out[len(want):, 2]))                     # padding points fail the kernel's clip
        got[isa] = out
    for isa in (1, 2, -1, "2-unaligned"):
        assert np.array_equal(got[isa].view(np.uint32), got[0].view(np.uint32))


def test_staging_pool_ordering_contract_cpu():
    """tests/cpp/stager_check.cpp: the pool (csrc/pcf_stager.hpp) with fake GPU hooks -- hand-over strictly in submission order
    for every mix of packers and raw lanes, each cloud exactly once with its own packed payload, drop_queued / drain semantics."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "high-fidelity-pointcloud-fusion_b200"), "all"], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(root, "host"), "all"], check=True)
    for rep in range(3):
        r = subprocess.run([os.path.join(root, "tests", "_build", "stager_check")], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("OK ") == 7
