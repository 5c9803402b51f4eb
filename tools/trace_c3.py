"""PCF_TRACE breakdown of process() on C3 (1000 frames, 1000^3 grid): per-kernel durations of the second (warm) pass on stderr."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pcfusion_b200 as pcf
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
scene = synth.plate_sweep(1000)
g, npf = scene.grid, scene.points_per_frame
for trace in (0, 1):
    os.environ["PCF_TRACE"] = str(trace)
    fus = pcf.Fusion(g.box, g.res, max_frames=1 << 16, log_capacity_hint=n * npf)
    os.environ["PCF_TRACE"] = "0"
    for rep in range(2 if not trace else 1):
        for b in range(0, n, 125):
            pts, poses = synth.frames_on_device(scene, b, 125)
            fus.push_frames_device(pts, 125, npf, 4, poses, b)
            fus.sync(); del pts
        if trace:
            print("---- trace ----", file=sys.stderr, flush=True)
        fus.update(); tu = fus.timings()["update_ms"]
        nv = fus.extract_raw(); te = fus.timings()
        print("trace", trace, "rep", rep, "voxels", nv, "update", tu, te, flush=True)
        fus.clear()
    fus.close()
