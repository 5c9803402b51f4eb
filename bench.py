#!/usr/bin/env python
"""bench.py -- fused points/sec of the frame-integration hot path + process() latency.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line (rank 0).

Workload at N=1 = BASELINE.json configs[1]: "200-frame synthetic turntable scan, 1 mm voxels, 0.5 m box".
A step = one pass of the hot path over the whole 200-frame sequence:
    ingest (clip -> FP64 transform -> crop -> voxel index -> occupancy -> point log)   <- timed: value / ms_per_step
    update + extract (normals, cylinder scoring, compacted extraction) and clear       <- reported as process_ms
`value`  : input points (clipped ones included) / ingest time, clouds already resident in HBM.
`e2e`    : same metric through the C ABI with HOST (pinned) clouds: every frame's H2D copy and the read-back of
           the integration summary are inside the timed region.
N>1      : frames sharded in contiguous blocks, one process per GPU, no collective on the ingest path (weak
           scaling: every rank integrates its own 200-frame block); the process()-time grid merge is timed apart.
--impl reference : the reference's CPU implementation of the same path on the host cores (oracle/_ref = the
           reference's own OccupancyGrid.hpp compiled against shim headers; falls back to the oracle port).
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "turntable200: 200 x 640x480 organized clouds, sphere r=0.15 m, 1 mm voxels, 0.5 m box, two elevation rings"
N_FRAMES = 200
BATCH = 200           # frames per ingest launch on the HBM-resident path (library limit: 256)
CPU_SAMPLE_FRAMES = 20


def make_scene(n_frames=N_FRAMES, rank=0, world=1):
    import importlib
    synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")
    # one global sequence of world*n_frames frames; this rank owns the contiguous block [rank*n, (rank+1)*n)
    return synth.sphere_turntable(n_frames * world, rings=2), rank * n_frames


def gen_frames(scene, first, n, threads=None):
    threads = threads or min(16, os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(threads) as ex:
        out = list(ex.map(scene.frame, range(first, first + n)))
    pts = np.stack([o[0] for o in out])
    poses = np.stack([o[1] for o in out])
    return pts, poses


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1])); smax = float(p[2])
                except ValueError:
                    continue
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            top = sorted(sm)[len(sm) // 2:]   # the loaded half of the samples
            out = {"sm_mhz": statistics.median(top), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline_run(frames, poses, grid, max_frames=CPU_SAMPLE_FRAMES):
    """Time the reference's CPU path on a bounded sample: add_frame over the first frames, then update+download."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    kind, label = ("ref_timing", "reference") if O.available("ref_timing") else ("oracle", "port")
    if kind == "oracle":
        O.build()
    n = min(max_frames, len(frames))
    og = O.OracleGrid(grid.box, grid.res, grid.clip_zmin, grid.clip_zmax, reserve_hint=1000, kind=kind)
    t0 = time.perf_counter()
    kept = 0
    for i in range(n):
        kept += og.add_frame(frames[i], poses[i])
    t1 = time.perf_counter()
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)          # the reference prints progress lines to stdout
    try:
        og.update()
        nout = len(og.download())
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    t2 = time.perf_counter()
    og.close()
    pts = n * frames.shape[1]
    return {"value": pts / (t1 - t0), "unit": "points/s", "cores": 1, "kind": label,
            "sample": f"first {n} of {len(frames)} frames of the workload ({pts} input points, {kept} kept); "
                      f"grid path single-threaded as in the reference (OG.hpp:190-193 pragmas are commented out)",
            "ingest_s": t1 - t0, "process_ms": (t2 - t1) * 1e3, "extracted_voxels": nout,
            "host_cpus": os.cpu_count()}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene, first = make_scene()
    n = CPU_SAMPLE_FRAMES
    frames, poses = gen_frames(scene, first, n)
    runs = []
    for s in range(args.warmup + args.steps):
        r = cpu_baseline_run(frames, poses, scene.grid, n)
        if s >= args.warmup:
            runs.append(r)
    total_pts = sum(n * frames.shape[1] for _ in runs)
    total_s = sum(r["ingest_s"] for r in runs)
    val = total_pts / total_s
    base = dict(runs[-1]); base["value"] = val
    line = {"impl": "reference", "metric": "fused points/sec", "value": val, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / len(runs), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64 transform / f32 statistics", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": base["sample"]},
            "process_ms": statistics.mean(r["process_ms"] for r in runs),
            "cpu_baseline": base,
            "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import pcfusion_b200 as pcf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene, first = make_scene(N_FRAMES, rank, world)
    grid = scene.grid
    npf = scene.points_per_frame
    t_gen = time.perf_counter()
    frames, poses = gen_frames(scene, first, N_FRAMES)
    t_gen = time.perf_counter() - t_gen

    dev_frames = torch.from_numpy(frames).cuda(local)                 # HBM-resident clouds (983 MB > 126 MB L2)
    host_frames = torch.from_numpy(frames).pin_memory()               # pinned host clouds for the e2e leg
    fus = pcf.Fusion(grid.box, grid.res, grid.clip_zmin, grid.clip_zmax, device=local,
                     max_frames=max(1 << 16, N_FRAMES * world + 1), log_capacity_hint=N_FRAMES * npf)
    stream = torch.cuda.ExternalStream(fus.stream, device=local)
    points_per_step = N_FRAMES * npf

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ingest_device():
        for b in range(0, N_FRAMES, BATCH):
            nb = min(BATCH, N_FRAMES - b)
            fus.push_frames_device(dev_frames[b], nb, npf, 4, poses[b:b + nb], first + b)

    sh = peer = None
    if world > 1:
        import importlib
        sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
        peer = sh.PeerExchange(fus)      # receive buffers mapped into every peer (CUDA IPC): the exchange kernel stores over NVLink

    def process_and_clear(keep=None):
        if world > 1:     # process() across ranks: slab-routed records written straight into the peers' buffers, then slab work
            n_local, _, tm = sh.merge_and_extract_v2(fus, peer=peer, gather_to=None)     # every rank keeps its own x-slab
            fus.clear()
            if keep is not None:
                nv = torch.tensor([n_local], dtype=torch.int64, device=f"cuda:{local}")
                dist.all_reduce(nv)
                keep.append({"update_ms": tm["exchange_ms"], "extract_device_ms": tm["slab_process_ms"], "extract_d2h_ms": 0.0,
                             "voxels": int(nv.item())})
            return
        fus.update()
        t = fus.timings()
        n = fus.extract_raw()
        t2 = fus.timings()
        fus.clear()
        if keep is not None:
            keep.append({"update_ms": t["update_ms"], "extract_device_ms": t2["extract_device_ms"],
                         "extract_d2h_ms": t2["extract_d2h_ms"], "voxels": n})

    # ---- value: HBM-resident ingest ------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()        # sampled from warm-up to the end of the e2e leg: the timed regions alone last only milliseconds
    for _ in range(args.warmup):
        ingest_device(); process_and_clear()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    proc = []
    fus.reset_stats()
    kept = 0
    t_wall = time.perf_counter()
    flush = torch.zeros(512 << 20, dtype=torch.int32, device=f"cuda:{local}")     # 2 GB >> 126 MB L2; ~0.35 ms of device work
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flush.sum()               # read-only sweep of 2 GB on the context's stream: replaces the L2 contents with clean
        ev[s][0].record(stream)       # lines and keeps the GPU busy while the host prepares the launch, so the event pair
                                      # brackets device work only (no host launch gap, no dirty-line write-back inside)
        ingest_device()
        ev[s][1].record(stream)
        if s == 0:
            kept = fus.count_kept()
        process_and_clear(proc)
    barrier()
    t_wall = time.perf_counter() - t_wall
    st = fus.stats()
    ingest_ms = [a.elapsed_time(b) for a, b in ev]
    total_ingest_ms = sum(ingest_ms)
    t = torch.tensor([total_ingest_ms], dtype=torch.float64, device=f"cuda:{local}")
    per_rank_ms = [float(t.item())]
    if world > 1:
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank_ms = [float(x.item()) for x in allt]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ingest_ms = float(t.item())
    value = points_per_step * world * args.steps / (max_ingest_ms * 1e-3)

    # ---- e2e: host (pinned) clouds through the C ABI, copies inside the timed region -------------------------
    def ingest_host():
        for i in range(N_FRAMES):
            fus.push_frame(host_frames[i], poses[i], first + i)
        return fus.count_kept()              # drain + D2H read of the integration summary

    for _ in range(max(1, min(args.warmup, 2))):
        ingest_host(); process_and_clear()
    e2e_s = 0.0
    e2e_proc = []
    barrier()
    for s in range(args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ingest_host()
        e2e_s += time.perf_counter() - t0
        t0 = time.perf_counter()
        process_and_clear()
        e2e_proc.append((time.perf_counter() - t0) * 1e3)
    barrier()
    t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = points_per_step * world * args.steps / float(t.item())
    # informational: the same leg with packed xyz clouds (12 B/point instead of the float4 layout's 16): PCIe is the bound
    e2e_packed = None
    if world == 1:
        host3 = torch.from_numpy(np.ascontiguousarray(frames[:, :, :3])).pin_memory()
        def ingest_host3():
            for i in range(N_FRAMES):
                fus.push_frame(host3[i], poses[i], first + i)
            return fus.count_kept()
        ingest_host3(); process_and_clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(args.steps):
            ingest_host3()
            if s + 1 < args.steps:
                fus.clear()
        e2e_packed = {"value": points_per_step * args.steps / (time.perf_counter() - t0), "unit": "points/s",
                      "h2d_bytes_per_step": int(points_per_step * 12), "note": "packed xyz host clouds (stride 3); clear() between steps inside the timed region"}
        process_and_clear()
        del host3
    clocks = sampler.stop()

    # ---- roofline of the dominant kernel (k_ingest) -----------------------------------------------------------
    launches_per_step = (N_FRAMES + BATCH - 1) // BATCH
    alg_bytes_step = points_per_step * 16 + kept * 20        # 16 B read / input point; 16 B record + 4 B grid probe / kept point
    per_launch_ms = statistics.mean(ingest_ms) / launches_per_step
    peak, peak_src = measured_peak()
    achieved = alg_bytes_step / launches_per_step / (per_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ingest_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "k_ingest_bulk<16>", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_step / launches_per_step, "launch_ms": per_launch_ms,
                "kept_fraction": kept / points_per_step}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_run(frames, poses, grid)

    if rank == 0:
        line = {
            "metric": "fused points/sec", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ingest_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 transform / f32 statistics", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": N_FRAMES, "points_per_frame": npf, "batch_frames_per_launch": BATCH,
                       "l2": "read-only sweep of 2 GB on the same stream before every timed step + inputs (983 MB per step) larger than the 126 MB L2",
                       "sharding": f"frames x{world}"},
            "process_ms": statistics.mean(p["update_ms"] + p["extract_device_ms"] + p["extract_d2h_ms"] for p in proc),
            "process_detail": dict({k: statistics.mean(p[k] for p in proc) for k in ("update_ms", "extract_device_ms", "extract_d2h_ms", "voxels")},
                                   note=("update_ms = exchange (plane histogram + viewpoint all-reduce, count all-gather, ONE compaction+peer-store "
                                         "kernel over NVLink, barrier); extract_device_ms = install + slab update + extract incl. D2H; "
                                         "rank 0's view") if world > 1 else "single GPU"),
            "step_wall_ms": 1e3 * t_wall / args.steps,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": int(points_per_step * 16),
                    "d2h_bytes_per_step": 4, "process_wall_ms": statistics.mean(e2e_proc)},
            "e2e_packed_xyz": e2e_packed,
            "ingest_ms_per_rank": [x / args.steps for x in per_rank_ms],
            "gpu_launches": int(st["kernel_launches"]),
            "clocks": clocks, "roofline": roofline, "gen_s": t_gen,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    fus.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
