"""torchrun --nproc-per-node N tools/multi_gpu_c4.py [frames] [update_every] : BASELINE config C4 (1920x1080 clouds, 0.5 m box @ 0.5 mm
= 1000^3 cells, an update pass after every 10 frames) with the frames of every round split over N real GPUs (replicated-state
mode, sharded.InterleavedSharded).  Rank 0 also runs the same schedule alone; the sharded extraction must have the same voxel
count and checksums.  Prints one JSON line."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
import pcfusion_b200 as pcf
sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 50
every = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
scene = synth.hires_sphere(n_frames)
g, npf = scene.grid, scene.points_per_frame


def run(fus, world_, rank_, il):
    ingest_ms, upd_ms = 0.0, 0.0
    stream = torch.cuda.ExternalStream(fus.stream, device=local)
    for start in range(0, n_frames, every):
        stop = min(start + every, n_frames)
        lo, hi = sh.frame_block(stop - start, rank_, world_)
        if hi > lo:
            pts, poses = synth.frames_on_device(scene, start + lo, hi - lo, dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fus.push_frames_device(pts, hi - lo, npf, 4, poses, start + lo)
            e1.record(stream)
            fus.sync()
            ingest_ms += e0.elapsed_time(e1)
            del pts
        t0 = time.perf_counter()
        if il is not None:
            il.update(start, stop)
        else:
            fus.update()
        torch.cuda.synchronize()
        upd_ms += (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    if il is not None:
        il.update(n_frames, n_frames)
        local_res, _ = il.extract(gather_to=None)          # every rank keeps its slab: no result traffic inside the timing
    else:
        fus.update()
        local_res = fus.extract()
    torch.cuda.synchronize()
    return ingest_ms, upd_ms, (time.perf_counter() - t0) * 1e3, local_res


fus = pcf.Fusion(g.box, g.res, device=local, max_frames=1 << 16, log_capacity_hint=n_frames * npf)
il = sh.InterleavedSharded(fus)
warm = torch.zeros(1, device=dev)
dist.all_reduce(warm)                                  # NCCL communicator set-up happens at the first collective: not part of the timings
run(fus, world, rank, il)                              # first pass: buffers grow to their final size
fus.clear()
il.bounds = None
dist.barrier()
ingest_ms, upd_ms, final_ms, res = run(fus, world, rank, il)
cs = bench.checksums(res)
t = torch.tensor([ingest_ms, upd_ms, final_ms], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
sums = torch.tensor(cs[:3], dtype=torch.int64, device=dev)
dist.all_reduce(sums)
xors = torch.tensor(cs[3:], dtype=torch.int64, device=dev)
allx = [torch.zeros_like(xors) for _ in range(world)]
dist.all_gather(allx, xors)
got = [int(sums[0]), int(sums[1]) % bench.MOD, int(sums[2])] + [int(np.bitwise_xor.reduce(np.array([int(a[i]) for a in allx], dtype=np.int64))) for i in range(5)]
line = {"config": f"C4 hires{n_frames}, update every {every} frames, rounds split x{world} (replicated state)", "points": n_frames * npf,
        "ingest_ms_max_rank": float(t[0]), "ingest_points_per_s": n_frames * npf / (float(t[0]) * 1e-3),
        "interleaved_updates_wall_ms": float(t[1]), "final_update_extract_wall_ms": float(t[2]), "extracted_voxels": got[0]}
torch.cuda.synchronize()
del il
fus.close()
if rank == 0:
    one = pcf.Fusion(g.box, g.res, device=local, max_frames=1 << 16, log_capacity_hint=n_frames * npf)
    run(one, 1, 0, None)
    one.clear()
    i1, u1, f1, r1 = run(one, 1, 0, None)
    want = bench.checksums(r1)
    want[1] %= bench.MOD
    line.update({"single_gpu_ingest_points_per_s": n_frames * npf / (i1 * 1e-3), "single_gpu_interleaved_updates_wall_ms": u1,
                 "single_gpu_final_update_extract_wall_ms": f1, "checksums_equal": [int(a) for a in got] == [int(b) for b in want]})
    one.close()
    print(json.dumps(line), flush=True)
dist.barrier()
dist.destroy_process_group()
