"""Shared comparison helpers: the parity bar is bit-exact for every integer AND float field, because the CUDA
path executes the oracle's float operations in the same order (no FMA); the north-star tolerances
(1e-5 m, 1e-4 rad) are asserted as well so a future relaxation of bit-exactness still has a stated bound."""
import numpy as np

RESULT_FIELDS = ["hash", "centroid", "normal", "sd", "mean_dist", "sd_dist", "count"]
STATE_FIELDS = ["hash", "buffer_len", "normal_found", "count", "normal", "viewpoint"]


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def assert_same(obj_a, obj_b, fields, what=""):
    for f in fields:
        x, y = getattr(obj_a, f), getattr(obj_b, f)
        assert x.shape == y.shape, f"{what}{f}: shape {x.shape} vs {y.shape}"
        if not bits_equal(x, y):
            xv, yv = x.reshape(len(x), -1), y.reshape(len(y), -1)
            bad = np.nonzero((xv.view(np.uint8) != yv.view(np.uint8)).any(axis=1))[0]
            raise AssertionError(f"{what}{f}: {len(bad)} of {len(x)} rows differ, first at {bad[0]}: {xv[bad[0]]} vs {yv[bad[0]]}")


def assert_result_parity(got, want, what=""):
    assert_same(got, want, RESULT_FIELDS, what)
    # stated tolerances of the north star (implied by bit equality; kept as the documented bound)
    if len(want.hash):
        assert np.nanmax(np.abs(got.centroid - want.centroid), initial=0) <= 1e-5
        a, b = got.normal.astype(np.float64), want.normal.astype(np.float64)
        ok = np.isfinite(a).all(axis=1) & np.isfinite(b).all(axis=1)
        ang = np.arctan2(np.linalg.norm(np.cross(a[ok], b[ok]), axis=1), np.sum(a[ok] * b[ok], axis=1))
        assert np.all(ang <= 1e-4)


def run_schedule(grid, scene, update_every=None, frames=None, device_push=None):
    """frames -> (update every k frames) -> final update.  `grid` is an OracleGrid or a pcfusion Fusion."""
    n = scene.n_frames if frames is None else frames
    for i in range(n):
        pts, T = scene.frame(i)
        if hasattr(grid, "push_frame"):
            grid.push_frame(pts, T, i)
        else:
            grid.add_frame(pts, T)
        if update_every and (i + 1) % update_every == 0:
            grid.update()
    grid.update()
