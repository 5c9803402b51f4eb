"""A/B of the k_score variants on C2 (200 frames): PCF_SCORE_UNR x PCF_SCORE_BALANCE -> process() device times; then one
PCF_TRACE pass (per-kernel durations on stderr) of the default configuration.  Run under gpurun."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import pcfusion_b200 as pcf

scene, _ = bench.make_scene()
frames, poses = bench.gen_frames(scene, 0, 200)
g, npf = scene.grid, scene.points_per_frame
dev = torch.from_numpy(frames).cuda()
ref = None


def run(env, trace=False):
    global ref
    for k in ("PCF_SCORE_UNR", "PCF_SCORE_BALANCE", "PCF_SCORE_COOP", "PCF_COOP_SLOTS", "PCF_TRACE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    fus = pcf.Fusion(g.box, g.res, log_capacity_hint=200 * npf)
    best = None
    for rep in range(3 if not trace else 2):
        fus.push_frames_device(dev, 200, npf, 4, poses, 0)
        fus.update(); tu = fus.timings()["update_ms"]
        n = fus.extract_raw(); te = fus.timings()
        if best is None or te["extract_device_ms"] < best["extract_device_ms"]:
            best = dict(te, update_ms=tu, voxels=n)
        if rep == 0:
            r = fus.extract()
            sig = (len(r), int(r.count.sum()), int(np.bitwise_xor.reduce(r.centroid.reshape(-1).view(np.uint32))))
            if ref is None:
                ref = sig
            best_same = sig == ref
        fus.clear()
    fus.close()
    return dict(env=env, same_result=best_same, **best)


for coop, bal, slots in ((0, 0, 16), (0, 1, 16), (1, 0, 8), (1, 1, 16), (1, 1, 8)):
    print(json.dumps(run({"PCF_SCORE_COOP": str(coop), "PCF_SCORE_BALANCE": str(bal), "PCF_COOP_SLOTS": str(slots)})), flush=True)
print("---- trace of the default configuration ----", file=sys.stderr, flush=True)
print(json.dumps(run({"PCF_TRACE": "1"}, trace=True)), flush=True)
