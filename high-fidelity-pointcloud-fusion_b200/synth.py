"""Seeded synthetic depth-cloud sequences for the offline replay driver.

The reference consumes `sensor_msgs/PointCloud2` clouds in the CAMERA frame plus a fusion<-camera pose
from tf (node.cpp:327-349).  There are no recorded bags, so every benchmark/test sequence is generated:
a pinhole depth camera looks at an analytic surface; each pixel's ray is intersected with the surface,
N(0, sigma) noise is added along the ray, and pixels that miss are NaN (what a RealSense emits).

Frames are float32 [H*W, 4] (x, y, z, pad) + float64 4x4 row-major pose; `seed = 1234 + frame_idx`.
All generation is numpy on the host and is never inside a timed region.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np


@dataclasses.dataclass
class GridSpec:
    box: tuple  # xmin, xmax, ymin, ymax, zmin, zmax (launch:8 order)
    res: float
    clip_zmin: float = 0.28  # node.cpp:92
    clip_zmax: float = 0.6   # node.cpp:93


def look_at_pose(cam_pos, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """fusion<-camera transform for an optical frame (z forward, x right, y down)."""
    c = np.asarray(cam_pos, dtype=np.float64)
    z = np.asarray(target, dtype=np.float64) - c
    z /= np.linalg.norm(z)
    upv = np.asarray(up, dtype=np.float64)
    if abs(np.dot(upv, z)) > 0.999:
        upv = np.array([0.0, 1.0, 0.0])
    x = np.cross(z, upv)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    T = np.eye(4)
    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, y, z, c
    return T


class Scene:
    """A sequence of (cloud, pose) frames over an analytic surface."""

    def __init__(self, name, grid: GridSpec, width, height, fx, poses, surface, noise_sigma=0.0003):
        self.name, self.grid, self.width, self.height, self.fx = name, grid, width, height, fx
        self.poses = poses
        self.surface = surface
        self.noise_sigma = noise_sigma
        u = (np.arange(width, dtype=np.float64) - (width - 1) / 2.0) / fx
        v = (np.arange(height, dtype=np.float64) - (height - 1) / 2.0) / fx
        uu, vv = np.meshgrid(u, v)
        d = np.stack([uu, vv, np.ones_like(uu)], axis=-1).reshape(-1, 3)
        self._dirs = d / np.linalg.norm(d, axis=1, keepdims=True)

    @property
    def n_frames(self):
        return len(self.poses)

    @property
    def points_per_frame(self):
        return self.width * self.height

    def pose(self, i) -> np.ndarray:
        return np.ascontiguousarray(self.poses[i], dtype=np.float64)

    def frame(self, i):
        T = self.pose(i)
        R, c = T[:3, :3], T[:3, 3]
        dw = self._dirs @ R.T
        t = self.surface(c, dw)  # distance along each ray, NaN for a miss
        rng = np.random.default_rng(1234 + i)
        t = t + rng.normal(0.0, self.noise_sigma, size=t.shape)
        pts = np.zeros((self._dirs.shape[0], 4), dtype=np.float32)
        pts[:, :3] = (self._dirs * t[:, None]).astype(np.float32)
        return pts, T


def frames_on_device(scene: "Scene", first, n, device="cuda"):
    """The same camera model evaluated with torch on the GPU: [n, H*W, 4] float32 clouds that never touch the host,
    for the large configurations (C3 / C4) where numpy generation would dominate the run.  Noise comes from torch's
    generator (seed 1234 + frame_idx), so these clouds are NOT bit-identical to Scene.frame(); parity tests keep using
    the numpy path."""
    import torch
    dirs = torch.from_numpy(scene._dirs).to(device)
    out = torch.zeros((n, dirs.shape[0], 4), dtype=torch.float32, device=device)
    poses = np.stack([scene.pose(first + i) for i in range(n)])
    gen = torch.Generator(device=device)
    for i in range(n):
        T = torch.from_numpy(poses[i]).to(device)
        t = scene.surface.torch(T[:3, 3], dirs @ T[:3, :3].T)
        gen.manual_seed(1234 + first + i)
        t = t + torch.randn(t.shape, generator=gen, device=device, dtype=torch.float64) * scene.noise_sigma
        out[i, :, :3] = (dirs * t[:, None]).to(torch.float32)
    if out.is_cuda:
        torch.cuda.synchronize(out.device)     # libpcfusion runs on its own stream: the clouds must be complete before they are pushed
    return out, poses


def _sphere_surface(radius):
    def hit_torch(c, d):
        import torch
        b = d @ c
        disc = b * b - (c @ c - radius * radius)
        t = -b - torch.sqrt(disc.clamp_min(0))
        return torch.where((disc > 0) & (t > 0), t, torch.full_like(t, float("nan")))

    def hit(c, d):
        b = d @ c
        disc = b * b - (c @ c - radius * radius)
        t = np.full(d.shape[0], np.nan)
        ok = disc > 0
        t[ok] = -b[ok] - np.sqrt(disc[ok])
        t[t <= 0] = np.nan
        return t
    hit.torch = hit_torch
    return hit


def _plate_surface(half, amp, period):
    k = 2.0 * math.pi / period

    def f(x, y):
        return amp * np.sin(k * x) * np.cos(k * y)

    def hit_torch(c, d):
        import torch
        t = (0.0 - c[2]) / d[:, 2]
        for _ in range(40):
            x, y = c[0] + t * d[:, 0], c[1] + t * d[:, 1]
            t = (amp * torch.sin(k * x) * torch.cos(k * y) - c[2]) / d[:, 2]
        x, y = c[0] + t * d[:, 0], c[1] + t * d[:, 1]
        return torch.where((x.abs() <= half) & (y.abs() <= half) & (t > 0), t, torch.full_like(t, float("nan")))

    def hit(c, d):
        t = (0.0 - c[2]) / d[:, 2]
        for _ in range(40):
            x, y = c[0] + t * d[:, 0], c[1] + t * d[:, 1]
            t = (f(x, y) - c[2]) / d[:, 2]
        x, y = c[0] + t * d[:, 0], c[1] + t * d[:, 1]
        t = np.where((np.abs(x) <= half) & (np.abs(y) <= half) & (t > 0), t, np.nan)
        return t
    hit.torch = hit_torch
    return hit


def sphere_turntable(n_frames=20, width=640, height=480, res=0.001, fx=None, rings=1, radius=0.15,
                     standoff=0.45, box_half=0.25, noise_sigma=0.0003) -> Scene:
    """C1 (rings=1, 18 deg steps for 20 frames) / C2 (rings=2 at +-20 deg elevation, 200 frames)."""
    fx = fx if fx is not None else 600.0 * width / 640.0
    poses = []
    per_ring = max(1, n_frames // rings)
    for i in range(n_frames):
        ring = min(i // per_ring, rings - 1)
        az = 2.0 * math.pi * (i % per_ring) / per_ring
        el = 0.0 if rings == 1 else math.radians(20.0 if ring == 0 else -20.0)
        c = standoff * np.array([math.cos(el) * math.cos(az), math.cos(el) * math.sin(az), math.sin(el)])
        poses.append(look_at_pose(c, (0, 0, 0)))
    g = GridSpec((-box_half, box_half) * 3, res)
    return Scene(f"sphere_turntable{n_frames}", g, width, height, fx, poses, _sphere_surface(radius), noise_sigma)


def plate_sweep(n_frames=1000, width=640, height=480, res=0.001, cols=40, noise_sigma=0.0003) -> Scene:
    """C3: raster of downward-looking poses over a 0.8 m wavy plate in a 1 m box."""
    fx = 600.0 * width / 640.0
    rows = max(1, math.ceil(n_frames / cols))
    poses = []
    for i in range(n_frames):
        r, cidx = divmod(i, cols)
        if r % 2:
            cidx = cols - 1 - cidx  # boustrophedon, like a robot raster
        x = -0.35 + 0.7 * (cidx / max(1, cols - 1))
        y = -0.35 + 0.7 * (r / max(1, rows - 1))
        poses.append(look_at_pose((x, y, 0.4), (x, y, 0.0), up=(0, 1, 0)))
    g = GridSpec((-0.5, 0.5) * 3, res)
    return Scene(f"plate_sweep{n_frames}", g, width, height, fx, poses, _plate_surface(0.4, 0.05, 0.4), noise_sigma)


def hires_sphere(n_frames=50, res=0.0005) -> Scene:
    """C4: the C1 scene at 1920x1080 and 0.5 mm voxels."""
    s = sphere_turntable(n_frames, 1920, 1080, res, fx=1800.0)
    s.name = f"hires_sphere{n_frames}"
    return s


def small_sphere(n_frames=6, width=160, height=120, res=0.005, noise_sigma=0.0006) -> Scene:
    """A seconds-scale scene for parity tests (coarse voxels so that sparse pixels still form a surface)."""
    s = sphere_turntable(n_frames, width, height, res, noise_sigma=noise_sigma)
    s.name = f"small_sphere{n_frames}"
    return s


def wavy_sheets_world(n_sheets=4, n_side=200, res=0.001, box_half=0.5, pts_per_voxel=(1, 4), seed=7):
    """C5-style state: world-frame points forming stacked one-voxel-thick wavy sheets (no camera).

    Returns (GridSpec, [ (xyz float32 [n,3], viewpoint float32[3]) per sheet ]).
    """
    g = GridSpec((-box_half, box_half) * 3, res)
    rng = np.random.default_rng(seed)
    out = []
    span = n_side * res
    for s in range(n_sheets):
        ix, iy = np.meshgrid(np.arange(n_side), np.arange(n_side), indexing="ij")
        x = -span / 2 + (ix + 0.5) * res
        y = -span / 2 + (iy + 0.5) * res
        z0 = -box_half * 0.8 + (s + 0.5) * (1.6 * box_half / n_sheets)
        z = z0 + 0.004 * np.sin(2 * math.pi * x / 0.1 + s) * np.cos(2 * math.pi * y / 0.13)
        reps = rng.integers(pts_per_voxel[0], pts_per_voxel[1] + 1, size=x.shape)
        xr, yr, zr = np.repeat(x.ravel(), reps.ravel()), np.repeat(y.ravel(), reps.ravel()), np.repeat(z.ravel(), reps.ravel())
        jit = rng.uniform(-0.45 * res, 0.45 * res, size=(xr.size, 3))
        pts = (np.stack([xr, yr, zr], axis=1) + jit * np.array([1, 1, 0.3])).astype(np.float32)
        out.append((pts, np.array([0.0, 0.0, box_half * 2], dtype=np.float32)))
    return g, out


def write_sequence(scene: Scene, path, n_frames=None, stride=4):
    """Record `scene` as a PCFSEQ1 file (host/sequence.hpp) for the C++ replay driver: the offline stand-in for a bag
    of `input_point_cloud` messages + tf poses (node.cpp:327-349)."""
    import struct
    n = scene.n_frames if n_frames is None else n_frames
    g = scene.grid
    res = np.float32(g.res)
    with open(path, "wb") as f:
        f.write(struct.pack("<8sIIII6d3ff2d", b"PCFSEQ1\0", n, scene.points_per_frame, stride, 0, *[float(b) for b in g.box],
                            res, res, res, 0.0, g.clip_zmin, g.clip_zmax))
        for i in range(n):
            pts, T = scene.frame(i)
            f.write(np.ascontiguousarray(T, np.float64).tobytes())
            f.write(np.ascontiguousarray(pts[:, :stride], np.float32).tobytes())
    return path
