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


# ---- exchange v2 (slab-routed records): routing arithmetic over gloo, world_size 2 --------------------------------
def _worker_v2(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh, grid, vps, log, n1 = _rank_state(rank, world)
    plane_cells = n1[1] * n1[2]
    x = (log[:, 3].view(np.uint32) // plane_cells).astype(np.int64)
    plane = torch.from_numpy(np.bincount(x, minlength=n1[0]).astype(np.int64))
    dist.all_reduce(plane)                                      # collective 1: the plane histogram
    bounds = sh.slab_bounds_from_points(plane.numpy(), world)
    halo = 3
    sel = [(x >= max(bounds[d] - halo, 0)) & (x < min(bounds[d + 1] + halo, n1[0])) for d in range(world)]
    mine = torch.tensor([int(s.sum()) for s in sel], dtype=torch.int64)
    rows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine)                                 # collective 2: the count matrix
    off, total = sh.route_offsets(torch.stack(rows).numpy())
    # the transfer itself (peer stores on the GPU) modelled as every rank contributing its block at its offset
    recv = [torch.zeros((int(total[d]), 4), dtype=torch.float32) for d in range(world)]
    for d in range(world):
        recv[d][off[rank][d]: off[rank][d] + int(mine[d])] = torch.from_numpy(log[sel[d]])
        dist.all_reduce(recv[d])                                # disjoint blocks: a sum assembles the buffer
    np.savez(os.path.join(out_dir, f"v2r{rank}.npz"), bounds=np.array(bounds), recv=recv[rank].numpy(), total=total)
    dist.destroy_process_group()


def test_exchange_v2_routing_world2(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_v2, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sh, grid1, vps1, log1, n1 = _rank_state(0, 1)
    x1 = (log1[:, 3].view(np.uint32) // (n1[1] * n1[2])).astype(np.int64)
    outs = [np.load(tmp_path / f"v2r{r}.npz") for r in range(world)]
    b = outs[0]["bounds"]
    assert np.array_equal(b, outs[1]["bounds"]) and b[0] == 0 and b[-1] == n1[0]
    for d in range(world):
        want = log1[(x1 >= max(b[d] - 3, 0)) & (x1 < min(b[d + 1] + 3, n1[0]))]     # slab + halo of the 1-rank log, arrival order
        assert np.array_equal(outs[d]["recv"].view(np.uint32), want.view(np.uint32))
    loads = [int(((x1 >= b[d]) & (x1 < b[d + 1])).sum()) for d in range(world)]
    assert sum(loads) == len(log1) and min(loads) > 0.3 * len(log1)


# ---- interleaved schedules across ranks (replicated state): the variable-size gathers over gloo, world_size 2 ------------
def _worker_rounds(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh, _, _, _, _ = _rank_state(rank, world)
    import golden_util as G
    import oracle as O
    fx = G.load("sphere_5mm")
    nf, every = len(fx.frames), 2
    merged_rounds, vp_table = [], torch.zeros((16, 4), dtype=torch.float32)
    for start in range(0, nf, every):
        stop = min(start + every, nf)
        lo, hi = sh.frame_block(stop - start, rank, world)          # this rank's frames of the round
        recs = []
        for i in range(start + lo, start + hi):
            pts, T = fx.frames[i], fx.poses[i]
            z = pts[:, 2].astype(np.float64)
            w = O.kat_transform(T, pts[(z < fx.clip[1]) & (z > fx.clip[0])])
            r = np.zeros((len(w), 4), np.float32)
            r[:, :3] = w
            r[:, 3] = np.uint32(i).view(np.float32)                 # (x, y, z, frame_idx): what pcf_round_export emits
            recs.append(r)
            vp_table[i] = torch.tensor([np.float32(T[0, 3]), np.float32(T[1, 3]), np.float32(T[2, 3]), 1.0])
        mine = torch.from_numpy(np.concatenate(recs) if recs else np.zeros((0, 4), np.float32))
        merged_rounds.append(sh._gather_var(mine, None).numpy())   # rank order == frame order == arrival order
        dist.all_reduce(vp_table[start:stop])                       # rows of this round only
    np.savez(os.path.join(out_dir, f"il{rank}.npz"), merged=np.concatenate(merged_rounds), vps=vp_table.numpy())
    dist.destroy_process_group()


def test_interleaved_round_gathers_world2(tmp_path):
    """Every rank must end up with the same merged record stream, equal to the single-rank arrival order, and with every
    frame's viewpoint row exactly once (rows are summed per round, never twice)."""
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_rounds, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import golden_util as G
    import oracle as O
    fx = G.load("sphere_5mm")
    want, vps = [], np.zeros((16, 4), np.float32)
    for i in range(len(fx.frames)):
        pts, T = fx.frames[i], fx.poses[i]
        z = pts[:, 2].astype(np.float64)
        w = O.kat_transform(T, pts[(z < fx.clip[1]) & (z > fx.clip[0])])
        r = np.zeros((len(w), 4), np.float32)
        r[:, :3] = w
        r[:, 3] = np.uint32(i).view(np.float32)
        want.append(r)
        vps[i] = [np.float32(T[0, 3]), np.float32(T[1, 3]), np.float32(T[2, 3]), 1.0]
    want = np.concatenate(want)
    outs = [np.load(tmp_path / f"il{r}.npz") for r in range(world)]
    for o in outs:
        assert np.array_equal(o["merged"].view(np.uint32), want.view(np.uint32))
        assert np.array_equal(o["vps"], vps)
