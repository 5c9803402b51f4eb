"""Ingest launches of C1 (one 20-frame launch, 500^3 grid) and C4 (10-frame launches of 1920x1080 clouds, 0.5 mm voxels, 1000^3 grid)
for ncu: tools/gpu_prof_ingest.sh captures the LAST k_ingest_bulk launch of each config.  Prints the CUDA-event time per launch."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import pcfusion_b200 as pcf
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")

which = os.environ.get("PROF_CONFIG", "C1")
if which == "C1":
    scene, nf, reps = synth.sphere_turntable(20), 20, 3
else:
    scene, nf, reps = synth.hires_sphere(50), 10, 3
g, npf = scene.grid, scene.points_per_frame
fus = pcf.Fusion(g.box, g.res, log_capacity_hint=nf * npf * 2)
stream = torch.cuda.ExternalStream(fus.stream)
flush = torch.zeros(128 << 20, dtype=torch.int32, device="cuda")
out = []
for rep in range(reps):
    first = 0 if which == "C1" else rep * nf          # C4: three consecutive 10-frame launches (the grid fills up as in the real run)
    pts, poses = synth.frames_on_device(scene, first, nf)
    if which == "C1" and rep:
        fus.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        flush.sum()
    e0.record(stream)
    fus.push_frames_device(pts, nf, npf, 4, poses, first)
    e1.record(stream)
    fus.sync()
    kept = fus.count_kept()
    out.append({"config": which, "launch": rep, "frames": nf, "ms": e0.elapsed_time(e1), "points_per_s": nf * npf / (e0.elapsed_time(e1) * 1e-3), "kept_total": kept})
    del pts
for o in out:
    print(json.dumps(o), flush=True)
fus.close()
