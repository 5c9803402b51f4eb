"""torchrun --nproc-per-node N tools/multi_gpu_check.py : frame-sharded fusion on N real GPUs (exchange v2 over CUDA IPC /
NVLink peer stores) must give the same bytes as one GPU fed every frame.  Rank 0 prints the verdict."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import pcfusion_b200 as pcf
from helpers import RESULT_FIELDS, bits_equal
sh = importlib.import_module("high-fidelity-pointcloud-fusion_b200.sharded")
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for name, scene in [("small_sphere", synth.small_sphere(12)), ("sphere 640x480 1 mm", synth.sphere_turntable(16))]:
    g = scene.grid
    fus = pcf.Fusion(g.box, g.res, device=local)
    peer = sh.DeviceExchange(fus)
    lo, hi = sh.frame_block(scene.n_frames, rank, world)
    for i in range(lo, hi):
        fus.push_frame(*scene.frame(i), i)
    for rep in range(2):                     # twice: the second round reuses the mapped buffers
        _, full, tm = sh.merge_and_extract_v3(fus, peer=peer, gather_to=0)
        if rank == 0:
            one = pcf.Fusion(g.box, g.res, device=local)
            for i in range(scene.n_frames):
                one.push_frame(*scene.frame(i), i)
            one.update()
            want = one.extract()
            same = all(bits_equal(getattr(full, f), getattr(want, f)) for f in RESULT_FIELDS)
            ok &= same
            print(f"{name} x{world} rep {rep}: {len(full)} voxels, byte-identical to 1 GPU: {same}, timings {tm}", flush=True)
            one.close()
        fus.clear()
        for i in range(lo, hi):
            fus.push_frame(*scene.frame(i), i)
    fus.close()
if rank == 0:
    print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
