"""Frame-sharded multi-GPU fusion: one context per GPU, frames in contiguous frame_idx blocks per rank, no
collective while frames are integrated, ONE exchange at process().

The reference has a single grid behind a mutex (node.cpp:132,291-296); what makes sharding exact here is that the
per-voxel state built during integration is commutative:
    occupancy / first inserting frame  -> elementwise MIN of the dense first-frame grids
    per-voxel point buffers            -> union of the rank logs; rank order == frame order == arrival order
    viewpoint table                    -> disjoint per-frame rows (SUM of zero-initialised tables)
After the merge every rank owns an x-slab of voxels (balanced by occupied-voxel count), runs normal estimation,
scoring and extraction for it, and the slab results concatenated in rank order are the reference's x-major scan
(OG.hpp:463-465) -- byte-identical to a single-GPU run.

`merge_and_extract` works on a list of in-process ranks (N contexts on one or several GPUs of one process: used by
the tests) or, with `group=` a torch.distributed process group, on the local rank (one process per GPU, NCCL).
Interleaved update schedules across ranks are not supported yet (the library refuses, see pcf_log_replace).
"""
from __future__ import annotations

import numpy as np
import torch


class _DevArray:
    """Wrap a raw device pointer as a torch tensor (no copy) through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr, "version": 2}


def _as_tensor(ptr, shape, typestr, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def frame_block(n_frames, rank, world):
    """Contiguous block [lo, hi) of the global frame sequence owned by `rank`."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def choose_slabs(plane_counts, world):
    """x-plane boundaries b[0..world] with ~equal occupied voxels per slab; plane_counts[x] = voxels before plane x."""
    pc = np.asarray(plane_counts, dtype=np.int64)
    n_planes = len(pc) - 1
    total = int(pc[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        x = int(np.searchsorted(pc, target, side="left"))
        bounds.append(min(max(x, bounds[-1]), n_planes))
    bounds.append(n_planes)
    return bounds


def rank_views(fus, device):
    """(grid int32 view, viewpoint table float32 view, compact log float32 [P,4]) of one context, all on device."""
    fus.sync()
    gp, cells = fus.grid_buffer()
    grid = _as_tensor(gp, (cells,), "<i4", device)
    vp, nf = fus.viewpoint_table()
    vps = _as_tensor(vp, (nf, 4), "<f4", device)
    lp, n = fus.log_compact()
    log = _as_tensor(lp, (max(n, 1), 4), "<f4", device)[:n]
    return grid, vps, log


def _finish_rank(fus, merged_log, slab):
    fus.set_slab(slab[0], slab[1])
    fus.log_replace(merged_log, merged_log.shape[0])
    fus.update()
    return fus.extract()


def _concat_results(parts):
    out = parts[0]
    for f in ("hash", "centroid", "normal", "sd", "mean_dist", "sd_dist", "count"):
        setattr(out, f, np.concatenate([getattr(p, f) for p in parts]))
    return out


def merge_and_extract_local(ranks):
    """In-process emulation: `ranks` = list of Fusion contexts (rank order = frame-block order).  Returns the
    merged extraction (same bytes as one context fed all frames)."""
    devs = [torch.device("cuda", f.device_index) for f in ranks]
    views = [rank_views(f, d) for f, d in zip(ranks, devs)]
    d0 = devs[0]
    grid = views[0][0].clone()
    vps = views[0][1].clone()
    for g, v, _ in views[1:]:
        grid = torch.minimum(grid, g.to(d0))
        vps += v.to(d0)
    merged = torch.cat([lg.to(d0) for _, _, lg in views], dim=0).contiguous()
    for (g, v, _), d in zip(views, devs):
        g.copy_(grid.to(d))
        v.copy_(vps.to(d))
    bounds = choose_slabs(ranks[0].plane_counts(), len(ranks))
    parts = []
    for r, (f, d) in enumerate(zip(ranks, devs)):
        parts.append(_finish_rank(f, merged.to(d), (bounds[r], bounds[r + 1])))
    return _concat_results(parts)


def merge_exchange(grid, vps, log, group=None):
    """The three collectives of process(): in-place MIN on the grid, SUM on the viewpoint table, and an all-gather
    of the variable-length rank logs (rank order).  Works on CPU tensors with gloo and CUDA tensors with NCCL."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dist.all_reduce(grid, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(vps, op=dist.ReduceOp.SUM, group=group)
    n_local = torch.tensor([log.shape[0]], dtype=torch.int64, device=log.device)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    padded = torch.zeros((mx, 4), dtype=log.dtype, device=log.device)
    padded[: log.shape[0]] = log
    gathered = torch.empty((world * mx, 4), dtype=log.dtype, device=log.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    merged = torch.cat([gathered[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0).contiguous()
    return merged, sizes


def merge_and_extract(fus, group=None, gather_to=0):
    """One process per GPU: merge this rank's context with its peers and extract its slab.
    Returns (local slab result, merged result on rank `gather_to` else None, timings dict in ms)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", fus.device_index)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    grid, vps, log = rank_views(fus, dev)
    merged, _ = merge_exchange(grid, vps, log, group)
    ev[1].record()
    bounds = choose_slabs(fus.plane_counts(), world)
    local = _finish_rank(fus, merged, (bounds[rank], bounds[rank + 1]))
    ev[2].record()
    torch.cuda.synchronize(dev)
    timings = {"exchange_ms": ev[0].elapsed_time(ev[1]), "slab_process_ms": ev[1].elapsed_time(ev[2])}
    parts = [None] * world if rank == gather_to else None
    dist.gather_object(local, parts, dst=gather_to, group=group)
    full = _concat_results(parts) if rank == gather_to else None
    return local, full, timings
