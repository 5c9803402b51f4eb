"""Generates tests/golden/*.npz: small seeded input sequences + the reference results for them.

Run in the build container (where /root/reference is mounted): the expected outputs are produced by
oracle/_ref/libogref_ordered.so -- the reference's UNMODIFIED OccupancyGrid.hpp compiled against the shim headers
(pins D1/D2/D3, see oracle/ref_shim/ogref_driver.cpp) -- and the generation asserts that the oracle restatement
gives the same bits.  The GPU box has no /root/reference; it checks the oracle and the CUDA path against
these files.  Inputs are stored explicitly (not regenerated) so BLAS/libm differences between machines cannot matter.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import pcfusion_b200  # noqa: E402,F401

synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")
RES_F = ["hash", "centroid", "normal", "sd", "mean_dist", "sd_dist", "count"]
STATE_F = ["hash", "buffer_len", "normal_found", "count", "normal", "viewpoint"]


def run(kind, grid, frames, poses, every):
    og = O.OracleGrid(grid.box, grid.res, grid.clip_zmin, grid.clip_zmax, kind=kind)
    for i, (p, T) in enumerate(zip(frames, poses)):
        og.add_frame(p, T)
        if every and (i + 1) % every == 0:
            og.update()
    og.update()
    return og.download(), og.state(), og.dims


def same(a, b, fields):
    return all(np.array_equal(getattr(a, f).view(np.uint8), getattr(b, f).view(np.uint8)) for f in fields)


def make(name, scene, every_list, grid=None, pose_premul=None):
    """grid: override of the scene's GridSpec (e.g. anisotropic resolution, offset box); pose_premul: rigid motion applied to
    every pose after the clouds were generated (moves the whole scene in the fusion frame)."""
    frames = [np.ascontiguousarray(scene.frame(i)[0][:, :3]) for i in range(scene.n_frames)]
    poses = [scene.pose(i) if pose_premul is None else pose_premul @ scene.pose(i) for i in range(scene.n_frames)]
    g = grid or scene.grid
    out = {"box": np.array(g.box, np.float64), "res": np.asarray(g.res, np.float32), "clip": np.array([g.clip_zmin, g.clip_zmax]),
           "frames": np.stack(frames), "poses": np.stack(poses), "schedules": np.array(every_list, np.int32)}
    for every in every_list:
        r_ref, s_ref, dims = run("ref_ordered", g, frames, poses, every)
        r_ora, s_ora, dims2 = run("oracle", g, frames, poses, every)
        assert dims == dims2
        assert same(r_ref, r_ora, RES_F) and same(s_ref, s_ora, ["hash", "buffer_len", "normal_found", "count"]), (name, every)
        if every == 0:   # canonical schedule: independent of the work-list order, so the stock build must agree too
            r_st, _, _ = run("ref", g, frames, poses, every)
            assert same(r_st, r_ref, RES_F)
        out["dims"] = np.array(dims, np.int32)
        for f in RES_F:
            out[f"res{every}_{f}"] = getattr(r_ref, f)
        for f in STATE_F:
            out[f"state{every}_{f}"] = getattr(s_ref, f)
        print(name, "every", every, "voxels", len(r_ref), "occupied", len(s_ref.hash), "count sum", int(r_ref.count.sum()))
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    make("sphere_5mm", synth.sphere_turntable(5, 120, 90, 0.005, noise_sigma=0.0006), [0, 1, 2])
    # finer voxels, denser pixels: long per-voxel buffers, many in-cylinder points
    make("sphere_2mm", synth.sphere_turntable(4, 200, 150, 0.002, radius=0.06, standoff=0.36, box_half=0.1, noise_sigma=0.0003), [0, 2])
    # launch-file style grid: three different resolutions (the walk uses xres_ on every axis, OG.hpp:405), a box that is not
    # centred and whose z = 0 face cuts the (lifted) sphere; update after every frame and after every third frame
    lift = np.eye(4); lift[:3, 3] = [0.013, -0.008, 0.09]
    make("sphere_aniso", synth.sphere_turntable(6, 120, 90, 0.005, noise_sigma=0.0006), [0, 1, 3],
         grid=synth.GridSpec((-0.21, 0.24, -0.2, 0.26, 0.0, 0.23), (0.005, 0.004, 0.006)), pose_premul=lift)
