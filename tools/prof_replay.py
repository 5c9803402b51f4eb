"""Replay for profiling: the bench workload (200 frames, bench.BATCH frames per ingest launch) -> update -> extract, 3 times.
PROF_VARIANTS="A=1,B=2;C=3" runs the replay once per ';'-separated environment set (each in a fresh context), so one
ncu capture can compare library variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import pcfusion_b200 as pcf
n = int(os.environ.get("PROF_FRAMES", "200"))
reps = int(os.environ.get("PROF_REPS", "3"))
B = bench.BATCH
scene, first = bench.make_scene(200)
frames, poses = bench.gen_frames(scene, 0, n)
dev = torch.from_numpy(frames).cuda()
g = scene.grid
for variant in os.environ.get("PROF_VARIANTS", "").split(";"):
    saved = dict(os.environ)
    for kv in filter(None, variant.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    fus = pcf.Fusion(g.box, g.res, log_capacity_hint=n * scene.points_per_frame)
    for rep in range(reps):
        for b in range(0, n, B):
            fus.push_frames_device(dev[b], min(B, n - b), scene.points_per_frame, 4, poses[b:b + B], b)
        fus.update()
        nv = fus.extract_raw()
        st = fus.stats()
        print("variant", variant or "-", "rep", rep, "voxels", nv, "kept", st["points_kept"], "occupied", st["occupied_voxels"], "normals", st["normals_found"],
              fus.timings(), flush=True)
        fus.clear()
    fus.close()
    os.environ.clear(); os.environ.update(saved)
