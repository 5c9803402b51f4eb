import glob
import os
import types

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RES_F = ["hash", "centroid", "normal", "sd", "mean_dist", "sd_dist", "count"]
STATE_F = ["hash", "buffer_len", "normal_found", "count", "normal", "viewpoint"]


def fixtures():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    res = z["res"]
    res = float(res) if res.ndim == 0 else tuple(float(r) for r in res)
    fx = types.SimpleNamespace(name=name, box=tuple(z["box"]), res=res, clip=tuple(z["clip"]), frames=z["frames"],
                               poses=z["poses"], schedules=[int(s) for s in z["schedules"]], dims=tuple(int(d) for d in z["dims"]))
    fx.result = {s: types.SimpleNamespace(**{f: z[f"res{s}_{f}"] for f in RES_F}) for s in fx.schedules}
    fx.state = {s: types.SimpleNamespace(**{f: z[f"state{s}_{f}"] for f in STATE_F}) for s in fx.schedules}
    return fx


def cases():
    out = []
    for n in fixtures():
        for s in load(n).schedules:
            out.append((n, s))
    return out


def replay(grid, fx, every):
    """frames -> update every `every` frames (0 = never) -> final update; grid is an OracleGrid or a Fusion."""
    for i in range(len(fx.frames)):
        if hasattr(grid, "push_frame"):
            grid.push_frame(fx.frames[i], fx.poses[i], i)
        else:
            grid.add_frame(fx.frames[i], fx.poses[i])
        if every and (i + 1) % every == 0:
            grid.update()
    grid.update()
