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
Interleaved update schedules across ranks (an update pass between frames, like the node's 5 s cleanGrid timer) use the
replicated-state mode at the bottom of this file: records and normal records are all-gathered at every pass, so the grid
state is identical on all ranks; ingest, normal estimation, scoring and extraction stay sharded.
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
    # libpcfusion works on its own CUDA stream: everything torch produced for it (copies, cat, collectives) must be complete
    torch.cuda.synchronize(torch.device("cuda", fus.device_index))
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


# ---- exchange v2: slab-routed records, compaction fused with the peer write (pcfusion.h "exchange v2") -------------
def slab_bounds_from_points(plane_points, world):
    """x-plane boundaries b[0..world] with ~equal numbers of log records per slab; plane_points[x] = records in plane x
    (summed over ranks)."""
    pp = np.asarray(plane_points, dtype=np.int64)
    cum = np.concatenate([[0], np.cumsum(pp)])
    return choose_slabs(cum, world)


def route_offsets(count_matrix):
    """count_matrix[s][d] = records rank s sends to rank d.  Returns (offset[s][d] of s's block inside d's receive
    buffer -- source-rank order = frame order = arrival order -- , total[d])."""
    m = np.asarray(count_matrix, dtype=np.int64)
    off = np.zeros_like(m)
    off[1:] = np.cumsum(m, axis=0)[:-1]
    return off, m.sum(axis=0)


def merge_and_extract_local_v2(ranks):
    """In-process emulation of exchange v2: `ranks` = contexts in frame-block order (one or several GPUs of this process).
    Every context scatters straight into the other contexts' receive buffers (plain device pointers; across processes
    the same pointers come from pcf_ipc_open)."""
    world = len(ranks)
    devs = [torch.device("cuda", f.device_index) for f in ranks]
    for f in ranks:
        f.sync()
    plane = sum(f.plane_point_counts().astype(np.int64) for f in ranks)
    bounds = slab_bounds_from_points(plane, world)
    vps = None
    for f, d in zip(ranks, devs):
        vp, nf = f.viewpoint_table()
        v = _as_tensor(vp, (nf, 4), "<f4", d).to(devs[0])
        vps = v.clone() if vps is None else vps + v
    counts = np.stack([f.exchange_counts(bounds) for f in ranks]).astype(np.int64)
    off, total = route_offsets(counts)
    bufs = [f.recv_buffer(int(total[r])) for r, f in enumerate(ranks)]
    for s, f in enumerate(ranks):
        f.exchange_scatter(bufs, off[s])
    parts = []
    for r, (f, d) in enumerate(zip(ranks, devs)):
        vp, nf = f.viewpoint_table()
        _as_tensor(vp, (nf, 4), "<f4", d).copy_(vps.to(d))
        torch.cuda.synchronize(d)
        f.install_records(bufs[r], int(total[r]))
        f.set_slab(bounds[r], bounds[r + 1])
        f.update()
        parts.append(f.extract())
    return _concat_results(parts)


def merge_and_extract_local_v3(ranks):
    """In-process emulation of the device-resident exchange (merge_and_extract_v3): the same library calls -- histogram,
    device-side slab bounds + routing counts, asynchronous scatter, region-keeping install -- with the collectives replaced
    by plain tensor sums / copies between the contexts of this process.  Also returns the bounds the device chose."""
    world = len(ranks)
    devs = [torch.device("cuda", f.device_index) for f in ranks]
    hists = []
    for f, d in zip(ranks, devs):
        hp, n_planes = f.exchange_hist()
        f.sync()
        hists.append(_as_tensor(hp, (n_planes,), "<i8", d))
    total_hist = sum(h.to(devs[0]) for h in hists)                   # the all-reduce
    vps = None
    for f, d in zip(ranks, devs):
        vp, nf = f.viewpoint_table()
        v = _as_tensor(vp, (nf, 4), "<f4", d).to(devs[0])
        vps = v.clone() if vps is None else vps + v
    rows = []
    for r, (f, d, h) in enumerate(zip(ranks, devs, hists)):
        h.copy_(total_hist.to(d))
        vp, nf = f.viewpoint_table()
        _as_tensor(vp, (nf, 4), "<f4", d).copy_(vps.to(d))
        torch.cuda.synchronize(d)
        row = _as_tensor(f.exchange_plan(world, r), (2 * world + 1,), "<i8", d)
        f.sync()
        rows.append(row.cpu().numpy().copy())                         # the all-gather
    rows = np.stack(rows)
    off, total = route_offsets(rows[:, :world])
    bounds = [int(b) for b in rows[0, world:]]
    assert all([int(b) for b in rr[world:]] == bounds for rr in rows)
    bufs = [f.recv_buffer(int(total[r])) for r, f in enumerate(ranks)]
    for s, f in enumerate(ranks):
        f.exchange_scatter_async(bufs, off[s])
    for f in ranks:
        f.sync()                                                      # the barrier
    parts = []
    for r, f in enumerate(ranks):
        f.install_records(bufs[r], int(total[r]))
        f.set_slab(bounds[r], bounds[r + 1])
        f.update()
        parts.append(f.extract())
    return _concat_results(parts), bounds


class PeerExchange:
    """One process per GPU: receive buffers mapped into every peer with CUDA IPC, so that pcf_exchange_scatter's stores
    travel over NVLink.  Handles are re-exchanged only when a receive buffer had to grow."""

    def __init__(self, fus, group=None):
        import torch.distributed as dist
        self.fus, self.group = fus, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.ptrs, self.cap = None, -1

    def buffers(self, my_total):
        import torch.distributed as dist
        need = torch.tensor([int(my_total) > self.cap], dtype=torch.int32, device=torch.device("cuda", self.fus.device_index))
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        if self.ptrs is None or int(need.item()):
            self.fus.ipc_close_all()
            dist.barrier(group=self.group)     # no rank may free its exported buffer while a peer still has it mapped
            self.cap = max(int(my_total) + int(my_total) // 4, 1 << 16, self.cap)   # head-room: the next process() rarely re-maps
            mine = self.fus.recv_buffer(self.cap)
            handles = [None] * self.world
            dist.all_gather_object(handles, self.fus.ipc_export(), group=self.group)
            self.ptrs = [mine if r == self.rank else self.fus.ipc_open(handles[r]) for r in range(self.world)]
        return self.ptrs


class DeviceExchange(PeerExchange):
    """State of the device-resident exchange (merge_and_extract_v3): the IPC-mapped receive buffers, a pinned landing zone
    for the gathered rows, and the token of the stream-ordered barrier.  Receive-buffer capacities are tracked for EVERY
    rank with the same rule on every rank, so that all ranks agree on when to re-map without talking to each other."""

    def __init__(self, fus, group=None):
        super().__init__(fus, group)
        dev = torch.device("cuda", fus.device_index)
        self.rows_dev = torch.empty((self.world, 2 * self.world + 1), dtype=torch.int64, device=dev)
        self.token = torch.zeros(1, dtype=torch.int32, device=dev)
        self.caps = [-1] * self.world

    def buffers_for(self, totals):
        import torch.distributed as dist
        if self.ptrs is None or any(int(t) > c for t, c in zip(totals, self.caps)):
            self.caps = [max(int(t) + int(t) // 4, 1 << 16, c) for t, c in zip(totals, self.caps)]
            self.fus.sync()
            self.fus.ipc_close_all()
            dist.barrier(group=self.group)     # no rank may free its exported buffer while a peer still has it mapped
            mine = self.fus.recv_buffer(self.caps[self.rank])
            handles = [None] * self.world
            dist.all_gather_object(handles, self.fus.ipc_export(), group=self.group)
            self.ptrs = [mine if r == self.rank else self.fus.ipc_open(handles[r]) for r in range(self.world)]
        return self.ptrs


def merge_and_extract_v3(fus, group=None, gather_to=None, peer=None):
    """One process per GPU, device-resident exchange.  Every step is enqueued on the context's stream -- the plane histogram,
    its all-reduce, the slab bounds + routing counts (computed on the device), the all-gather of the rows, the ONE compaction
    kernel that stores into the peers' buffers over NVLink, a one-word all-reduce as the stream-ordered barrier, the install
    -- and the host waits exactly once, for the gathered R x (2R+1) rows that size the receive buffers.
    Returns (local slab result or voxel count, merged result on rank `gather_to` else None, timings dict in ms)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", fus.device_index)
    peer = peer or DeviceExchange(fus, group)
    stream = torch.cuda.ExternalStream(fus.stream, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.cuda.stream(stream):          # torch's collectives order themselves against the CURRENT stream = the library's
        ev[0].record()
        hp, n_planes = fus.exchange_hist()
        dist.all_reduce(_as_tensor(hp, (n_planes,), "<i8", dev), group=group)
        vp, nf = fus.viewpoint_table()
        dist.all_reduce(_as_tensor(vp, (nf, 4), "<f4", dev), group=group)      # disjoint per-frame rows: the sum is exact
        row = _as_tensor(fus.exchange_plan(world, rank), (2 * world + 1,), "<i8", dev)
        dist.all_gather_into_tensor(peer.rows_dev, row, group=group)
        # the one host wait of the exchange.  (A plain synchronous copy on purpose: a non_blocking copy into a pinned tensor
        # makes torch's host allocator remember this -- externally owned -- stream and record an event on it when the tensor
        # is freed, which may be after the context was destroyed.)
        rows = peer.rows_dev.cpu().numpy()
    off, total = route_offsets(rows[:, :world])
    bounds = [int(b) for b in rows[rank, world:]]
    ptrs = peer.buffers_for(total)
    with torch.cuda.stream(stream):
        fus.exchange_scatter_async(ptrs, off[rank])
        dist.all_reduce(peer.token, group=group)          # stream-ordered barrier: every rank's peer stores have landed
        ev[1].record()
    fus.install_records(ptrs[rank], int(total[rank]))
    fus.set_slab(bounds[rank], bounds[rank + 1])
    fus.update()
    local = fus.extract() if gather_to is not None else fus.extract_raw()
    with torch.cuda.stream(stream):
        ev[2].record()
    torch.cuda.synchronize(dev)
    timings = {"exchange_ms": ev[0].elapsed_time(ev[1]), "slab_process_ms": ev[1].elapsed_time(ev[2]),
               "records_in": int(total[rank]), "records_out": int(rows[rank, :world].sum())}
    if gather_to is None:                      # every rank keeps (or writes) its own slab: no result traffic at all
        return local, None, timings
    parts = [None] * world if rank == gather_to else None
    dist.gather_object(local, parts, dst=gather_to, group=group)
    full = _concat_results(parts) if rank == gather_to else None
    return local, full, timings


def merge_and_extract_v2(fus, group=None, gather_to=0, peer=None):
    """One process per GPU, exchange v2.  Collectives: two tiny all-reduces (plane histogram, viewpoint table), one
    all-gather of the R x R count matrix, then ONE kernel per rank that compacts and writes into the peers' buffers,
    and a barrier.  Returns (local slab result, merged result on rank `gather_to` else None, timings dict in ms)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", fus.device_index)
    peer = peer or PeerExchange(fus, group)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    fus.sync()
    plane = torch.from_numpy(fus.plane_point_counts().astype(np.int64)).to(dev)
    dist.all_reduce(plane, group=group)
    bounds = slab_bounds_from_points(plane.cpu().numpy(), world)
    vp, nf = fus.viewpoint_table()
    dist.all_reduce(_as_tensor(vp, (nf, 4), "<f4", dev), group=group)
    torch.cuda.synchronize(dev)               # the library reads the table on its own stream
    mine = torch.from_numpy(fus.exchange_counts(bounds).astype(np.int64)).to(dev)
    rows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine, group=group)
    off, total = route_offsets(torch.stack(rows).cpu().numpy())
    ptrs = peer.buffers(int(total[rank]))
    dist.barrier(group=group)                 # every peer has finished reading its receive buffer from the last round
    fus.exchange_scatter(ptrs, off[rank])
    dist.barrier(group=group)                 # every rank's stores have landed (pcf_exchange_scatter synchronises its stream)
    ev[1].record()
    fus.install_records(ptrs[rank], int(total[rank]))
    fus.set_slab(bounds[rank], bounds[rank + 1])
    fus.update()
    local = fus.extract() if gather_to is not None else fus.extract_raw()
    ev[2].record()
    torch.cuda.synchronize(dev)
    timings = {"exchange_ms": ev[0].elapsed_time(ev[1]), "slab_process_ms": ev[1].elapsed_time(ev[2]),
               "records_in": int(total[rank]), "records_out": int(mine.sum().item())}
    if gather_to is None:                      # every rank keeps (or writes) its own slab: no result traffic at all
        return local, None, timings
    parts = [None] * world if rank == gather_to else None
    dist.gather_object(local, parts, dst=gather_to, group=group)
    full = _concat_results(parts) if rank == gather_to else None
    return local, full, timings


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


# ---- interleaved update schedules across ranks: replicated grid state (pcfusion.h "interleaved update schedules") ------------
def _gather_var(t, group, emulate=None):
    """All-gather of tensors whose first dimension differs per rank, concatenated in rank order."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x.item()) for x in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * mx: r * mx + sizes[r]] for r in range(world)], dim=0).contiguous()


class InterleavedSharded:
    """One process per GPU.  Usage per round (the frames between two update passes, split over the ranks in contiguous
    sub-blocks with frame_block()):  push this rank's frames into `fus` as usual, then update(frame_lo, frame_hi) with the
    global frame range of the round; at the end extract()."""

    def __init__(self, fus, group=None):
        import torch.distributed as dist
        self.fus, self.group = fus, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.dev = torch.device("cuda", fus.device_index)
        self.bounds = None

    def exchange_round(self, frame_lo, frame_hi):
        import torch.distributed as dist
        p, n = self.fus.round_export()
        mine = _as_tensor(p, (max(n, 1), 4), "<f4", self.dev)[:n]
        merged = _gather_var(mine, self.group)
        vp, nf = self.fus.viewpoint_table()
        dist.all_reduce(_as_tensor(vp, (nf, 4), "<f4", self.dev)[frame_lo:frame_hi], group=self.group)   # rows of this round only
        torch.cuda.synchronize(self.dev)
        self.fus.round_install(merged, merged.shape[0])

    def _slabs(self):
        self.bounds = slab_bounds_from_points(self.fus.plane_point_counts(), self.world)   # the log is replicated: same bounds everywhere
        self.fus.set_slab(self.bounds[self.rank], self.bounds[self.rank + 1])

    def update(self, frame_lo, frame_hi):
        self.exchange_round(frame_lo, frame_hi)
        self._slabs()
        pc, pn, n = self.fus.update_local()
        cells = _gather_var(_as_tensor(pc, (max(n, 1),), "<i4", self.dev)[:n], self.group)
        nrms = _gather_var(_as_tensor(pn, (max(n, 1), 4), "<f4", self.dev)[:n], self.group)
        torch.cuda.synchronize(self.dev)
        self.fus.update_commit(cells, nrms, cells.shape[0])

    def extract(self, gather_to=0):
        """(this rank's slab result, merged result on rank `gather_to`); gather_to=None keeps every slab where it is."""
        import torch.distributed as dist
        if self.bounds is None:
            self._slabs()
        local = self.fus.extract()
        if gather_to is None:
            return local, None
        parts = [None] * self.world if self.rank == gather_to else None
        dist.gather_object(local, parts, dst=gather_to, group=self.group)
        return local, (_concat_results(parts) if self.rank == gather_to else None)


def interleaved_local(ranks, scene_frames, update_every):
    """In-process emulation of InterleavedSharded on a list of contexts (tests): frames = [(pts, pose)], an update pass after
    every `update_every` frames (and a final one); the gathers are plain tensor concatenations between the contexts."""
    world = len(ranks)
    devs = [torch.device("cuda", f.device_index) for f in ranks]
    n = len(scene_frames)
    bounds = None

    def do_update(lo, hi):
        nonlocal bounds
        recs, vps = [], None
        for f, d in zip(ranks, devs):
            p, k = f.round_export()
            recs.append(_as_tensor(p, (max(k, 1), 4), "<f4", d)[:k].to(devs[0]))
            vp, nf = f.viewpoint_table()
            v = _as_tensor(vp, (nf, 4), "<f4", d)[lo:hi].to(devs[0])
            vps = v.clone() if vps is None else vps + v
        merged = torch.cat(recs, dim=0).contiguous()
        for f, d in zip(ranks, devs):
            vp, nf = f.viewpoint_table()
            _as_tensor(vp, (nf, 4), "<f4", d)[lo:hi].copy_(vps.to(d))
            m = merged.to(d)
            torch.cuda.synchronize(d)
            f.round_install(m, m.shape[0])
        bounds = slab_bounds_from_points(ranks[0].plane_point_counts(), world)
        cells, nrms = [], []
        for r, (f, d) in enumerate(zip(ranks, devs)):
            f.set_slab(bounds[r], bounds[r + 1])
            pc, pn, k = f.update_local()
            cells.append(_as_tensor(pc, (max(k, 1),), "<i4", d)[:k].to(devs[0]).clone())
            nrms.append(_as_tensor(pn, (max(k, 1), 4), "<f4", d)[:k].to(devs[0]).clone())
        cells, nrms = torch.cat(cells).contiguous(), torch.cat(nrms, dim=0).contiguous()
        for f, d in zip(ranks, devs):
            cc, nn_ = cells.to(d), nrms.to(d)
            torch.cuda.synchronize(d)
            f.update_commit(cc, nn_, cc.shape[0])

    start = 0
    while start < n:
        stop = min(start + (update_every or n), n)
        for r, f in enumerate(ranks):
            lo, hi = frame_block(stop - start, r, world)
            for i in range(start + lo, start + hi):
                f.push_frame(scene_frames[i][0], scene_frames[i][1], i)
        do_update(start, stop)
        start = stop
    if update_every and n % update_every == 0:
        do_update(n, n)                                   # the final pass of run_schedule() (no new frames)
    return _concat_results([f.extract() for f in ranks])
