"""Replay for profiling: the bench workload (200 frames, bench.BATCH frames per ingest launch) -> update -> extract, 3 times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import pcfusion_b200 as pcf
n = int(os.environ.get("PROF_FRAMES", "200"))
B = bench.BATCH
scene, first = bench.make_scene(200)
frames, poses = bench.gen_frames(scene, 0, n)
dev = torch.from_numpy(frames).cuda()
g = scene.grid
fus = pcf.Fusion(g.box, g.res, log_capacity_hint=n * scene.points_per_frame)
for rep in range(3):
    for b in range(0, n, B):
        fus.push_frames_device(dev[b], min(B, n - b), scene.points_per_frame, 4, poses[b:b + B], b)
    fus.update()
    nv = fus.extract_raw()
    print("rep", rep, "voxels", nv, fus.timings(), flush=True)
    fus.clear()
fus.close()
