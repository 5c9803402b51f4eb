"""torchrun --nproc-per-node N tools/multi_gpu_c3.py [frames] : BASELINE config C3 (plate sweep, 1 m box, 1 mm voxels = 1000^3
cells) frame-sharded over N real GPUs.  Every rank generates and integrates its contiguous block of frames, then the
ranks run process() with exchange v2 (peer stores over NVLink).  Rank 0 additionally integrates ALL frames in a second
context; the sharded extraction must have the same voxel count and the same checksums (sum of hashes, sum of counts,
XOR of the float bit patterns of every output field), i.e. it must be byte-identical up to the concatenation order,
which the strict x-major order of the slabs fixes.  Prints one JSON line."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import pcfusion_b200 as pcf
sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
scene = synth.plate_sweep(n_frames)
g, npf = scene.grid, scene.points_per_frame


MOD = 1 << 59        # per-rank sums are reduced mod 2^59 so that the int64 all-reduce over <= 8 ranks cannot overflow


def checksums(res):
    out = [len(res), int(res.hash.sum(dtype=np.uint64)) % MOD, int(res.count.sum(dtype=np.int64))]
    for f in ("centroid", "normal", "sd", "mean_dist", "sd_dist"):
        out.append(int(np.bitwise_xor.reduce(getattr(res, f).reshape(-1).view(np.uint32))))
    return out


def ingest(fus, lo, hi, batch=125):
    ms = 0.0
    stream = torch.cuda.ExternalStream(fus.stream)
    for b in range(lo, hi, batch):
        k = min(batch, hi - b)
        pts, poses = synth.frames_on_device(scene, b, k, dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fus.push_frames_device(pts, k, npf, 4, poses, b)
        e1.record(stream)
        fus.sync()
        ms += e0.elapsed_time(e1)
        del pts
    return ms


lo, hi = sh.frame_block(n_frames, rank, world)
fus = pcf.Fusion(g.box, g.res, device=local, max_frames=max(1 << 16, n_frames + 1), log_capacity_hint=(hi - lo) * npf)
peer = sh.DeviceExchange(fus)
line = {}
for rep in range(2):                       # second repetition = warm buffers
    ms = ingest(fus, lo, hi)
    dist.barrier()
    t0 = time.perf_counter()
    _, _, tm = sh.merge_and_extract_v3(fus, peer=peer, gather_to=None)     # every rank keeps its own x-slab
    dist.barrier()
    wall = (time.perf_counter() - t0) * 1e3
    if rep == 1:
        local_res = fus.extract()          # the slab's result again (extraction is idempotent), for the checksums only
        cs = torch.tensor(checksums(local_res)[:3], dtype=torch.int64, device=dev)      # counts and sums add up across slabs
        xs = torch.tensor(checksums(local_res)[3:], dtype=torch.int64, device=dev)
        dist.all_reduce(cs)
        allx = [torch.zeros_like(xs) for _ in range(world)]
        dist.all_gather(allx, xs)
        t = torch.tensor([ms, tm["exchange_ms"], tm["slab_process_ms"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line = {"config": f"C3 sweep{n_frames} frame-sharded x{world}", "frames_per_rank": hi - lo, "points": n_frames * npf,
                "ingest_ms_max_rank": float(t[0]), "ingest_points_per_s": n_frames * npf / (float(t[0]) * 1e-3),
                "exchange_ms": float(t[1]), "slab_process_ms": float(t[2]), "process_wall_ms": wall,
                "extracted_voxels": int(cs[0]), "records_in_rank0": tm["records_in"]}
        sharded_sums = [int(cs[0]), int(cs[1]) % MOD, int(cs[2])] + [int(np.bitwise_xor.reduce(np.array([int(a[i]) for a in allx], dtype=np.int64))) for i in range(5)]
    fus.clear()
fus.close()
if rank == 0:
    one = pcf.Fusion(g.box, g.res, device=local, max_frames=max(1 << 16, n_frames + 1), log_capacity_hint=n_frames * npf)
    ms1 = ingest(one, 0, n_frames)
    one.update()
    res = one.extract()
    want = checksums(res)
    line["single_gpu_ingest_points_per_s"] = n_frames * npf / (ms1 * 1e-3)
    line["single_gpu_voxels"] = want[0]
    line["checksums_equal"] = [int(a) for a in sharded_sums] == [int(b) for b in want]
    one.close()
    print(json.dumps(line), flush=True)
dist.barrier()
dist.destroy_process_group()
