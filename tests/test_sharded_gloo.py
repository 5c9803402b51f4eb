"""CPU, world_size 2, gloo: the host-side logic of the frame-sharded merge (sharded.py) -- frame blocks, slab choice
and the three collectives of process() -- checked against a numpy model of the commutative per-voxel state built
with the oracle's own voxel arithmetic."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMPTY = 0x7FFFFFFF


def _rank_state(rank, world):
    """Per-rank (grid, viewpoint table, log) exactly as a context would hold them, built with numpy + oracle KATs."""
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle as O
    import golden_util as G
    fx = G.load("sphere_5mm")
    og = O.OracleGrid(fx.box, fx.res, fx.clip[0], fx.clip[1])
    n1 = [d + 1 for d in og.dims]
    cells = n1[0] * n1[1] * n1[2]
    nf = len(fx.frames)
    import pcfusion_b200  # noqa: F401
    sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
    lo, hi = sh.frame_block(nf, rank, world)
    grid = np.full(cells, EMPTY, np.int32)
    vps = np.zeros((16, 4), np.float32)
    logs = []
    for i in range(lo, hi):
        pts, T = fx.frames[i], fx.poses[i]
        clip = (pts[:, 2].astype(np.float64) < fx.clip[1]) & (pts[:, 2].astype(np.float64) > fx.clip[0])
        w = O.kat_transform(T, pts[clip])
        ijk, valid = O.kat_voxel(og, w)
        w, ijk = w[valid == 1], ijk[valid == 1].astype(np.int64)
        cell = (ijk[:, 0] * n1[1] + ijk[:, 1]) * n1[2] + ijk[:, 2]
        np.minimum.at(grid, cell, i)
        vps[i] = [np.float32(T[0, 3]), np.float32(T[1, 3]), np.float32(T[2, 3]), 1.0]
        rec = np.zeros((len(w), 4), np.float32)
        rec[:, :3] = w
        rec[:, 3] = cell.astype(np.uint32).view(np.float32)
        logs.append(rec)
    log = np.concatenate(logs) if logs else np.zeros((0, 4), np.float32)
    return sh, grid, vps, log, n1


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh, grid, vps, log, n1 = _rank_state(rank, world)
    tg, tv, tl = torch.from_numpy(grid.copy()), torch.from_numpy(vps.copy()), torch.from_numpy(log.copy())
    merged, sizes = sh.merge_exchange(tg, tv, tl)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), grid=tg.numpy(), vps=tv.numpy(), merged=merged.numpy(), sizes=np.array(sizes))
    dist.destroy_process_group()


def test_merge_exchange_world2(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single-rank model of the same state
    sh, grid1, vps1, log1, n1 = _rank_state(0, 1)
    outs = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    for o in outs:
        assert np.array_equal(o["grid"], grid1)                     # MIN-merge == sequential first-frame
        assert np.array_equal(o["vps"], vps1)
        assert np.array_equal(o["merged"].view(np.uint32), log1.view(np.uint32))   # rank order == arrival order
    assert sum(outs[0]["sizes"]) == len(log1) and all(s > 0 for s in outs[0]["sizes"])


def test_frame_blocks_and_slabs():
    for p in (ROOT,):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pcfusion_b200  # noqa: F401
    sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
    for n, w in [(200, 8), (1000, 8), (7, 4), (3, 8)]:
        blocks = [sh.frame_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        assert max(b[1] - b[0] for b in blocks) - min(b[1] - b[0] for b in blocks) <= 1
    # cumulative voxel counts per x-plane -> monotone boundaries with balanced loads
    rng = np.random.default_rng(0)
    per_plane = rng.integers(0, 1000, 500)
    per_plane[:100] = 0
    pc = np.concatenate([[0], np.cumsum(per_plane)])
    for w in (1, 2, 4, 8):
        b = sh.choose_slabs(pc, w)
        assert b[0] == 0 and b[-1] == 500 and all(b[i] <= b[i + 1] for i in range(w))
        loads = [pc[b[i + 1]] - pc[b[i]] for i in range(w)]
        assert sum(loads) == pc[-1] and max(loads) - min(loads) <= 2 * per_plane.max()
