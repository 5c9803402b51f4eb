#!/usr/bin/env python
"""Run the BASELINE.json configurations C1..C5 on one B200 and print one JSON line each.

    python tools/run_configs.py C1 C2 C3 C4 C5 [--oracle]     (under gpurun; results also appended to gpurun_out/configs.jsonl)

Per config: HBM-resident ingest rate (points/s, CUDA events on the context stream), process() time (update + extract
device ms + D2H), occupied voxels, extracted voxels, and the size-independent checks (x-major order strictly
increasing, sum of buffer lengths == kept points, extraction idempotent, clear() empties).  --oracle adds a
bit-exact comparison with the CPU oracle where the oracle's dense grid fits comfortably (C1, C2)."""
import argparse, importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import pcfusion_b200 as pcf
synth = importlib.import_module("high-fidelity-pointcloud-fusion_b200.synth")


def checks(fus, res, kept, canonical=True):
    out = {}
    h = res.hash.astype(np.uint64)
    out["x_major_sorted"] = bool(np.all(h[1:] > h[:-1])) if len(h) > 1 else True
    if fus.stats()["occupied_voxels"] <= 20_000_000:
        st = fus.state()
        out["buffered_points"] = int(st.buffer_len.sum())
        out["normals_found"] = int(st.normal_found.sum())
        if canonical:      # every kept point is buffered when no update ran between frames
            out["buffer_sum_eq_kept"] = out["buffered_points"] == kept
        del st
    res2 = fus.extract()
    out["extract_idempotent"] = all(np.array_equal(getattr(res, f).view(np.uint8), getattr(res2, f).view(np.uint8))
                                    for f in ("hash", "centroid", "normal", "sd", "mean_dist", "sd_dist", "count"))
    return out


def run_scene(name, scene, n_frames, batch, update_every=0, oracle=False, frames_per_gen=50, passes=2):
    """passes=2: the first pass warms the allocator (scratch buffers grow on demand), the second is reported."""
    line = None
    for p in range(passes):
        line = _run_scene(name, scene, n_frames, batch, update_every, oracle and p == passes - 1, frames_per_gen, keep=line, numpy_frames=oracle)
    return line


_CTX = {}
_ORACLE_KIND = "oracle"    # --oracle-kind ref_ordered: the reference's own header (dense grid: 16 GB at 1000^3) instead of the restatement
_PREMUL = None     # optional fixed transform applied to every pose after the clouds were generated (layout experiments)


def _run_scene(name, scene, n_frames, batch, update_every, oracle, frames_per_gen, keep=None, numpy_frames=False):
    g = scene.grid
    npf = scene.points_per_frame
    if name not in _CTX:
        for k in [k for k in _CTX if k != "_flush"]:
            del _CTX[k]
        _CTX[name] = pcf.Fusion(g.box, g.res, g.clip_zmin, g.clip_zmax, max_frames=max(1 << 16, n_frames + 1), log_capacity_hint=n_frames * npf)
    fus = _CTX[name]
    stream = torch.cuda.ExternalStream(fus.stream)
    flush = _CTX.setdefault("_flush", torch.zeros(128 << 20, dtype=torch.int32, device="cuda"))
    ingest_ms, kept_pts, upd_ms = 0.0, 0, 0.0
    og = None
    if oracle:
        import oracle as O
        og = O.OracleGrid(g.box, g.res, g.clip_zmin, g.clip_zmax, kind=_ORACLE_KIND)
    t_cpu = 0.0
    done = 0
    while done < n_frames:
        nb = min(frames_per_gen, n_frames - done)
        if oracle or numpy_frames:
            frames, poses = bench.gen_frames(scene, done, nb)
            dev = torch.from_numpy(frames).cuda()
        else:
            dev, poses = synth.frames_on_device(scene, done, nb)
        if _PREMUL is not None:
            poses = np.stack([_PREMUL @ T for T in poses])
        b = 0
        while b < nb:
            k = min(batch, nb - b)
            if update_every:
                k = min(k, update_every - (done + b) % update_every)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                flush.sum()           # read-only 512 MB sweep: clean L2 + hides the host-side launch latency (as in bench.py)
            e0.record(stream)
            fus.push_frames_device(dev[b], k, npf, 4, poses[b:b + k], done + b)
            e1.record(stream)
            fus.sync()
            ingest_ms += e0.elapsed_time(e1)
            if og is not None:
                t0 = time.perf_counter()
                for i in range(b, b + k):
                    og.add_frame(frames[i], poses[i])
                t_cpu += time.perf_counter() - t0
            b += k
            if update_every and (done + b) % update_every == 0:
                fus.update(); upd_ms += fus.timings()["update_ms"]
                print(f"[{time.strftime('%X')}] {name}: {done + b} frames, update {fus.timings()['update_ms']:.2f} ms", file=sys.stderr, flush=True)
                if og is not None:
                    og.update()
        done += nb
        del dev
    kept_pts = fus.count_kept()
    print(f"[{time.strftime('%X')}] {name}: ingest done, kept {kept_pts}", file=sys.stderr, flush=True)
    fus.update(); t_u = fus.timings()["update_ms"]
    print(f"[{time.strftime('%X')}] {name}: final update {t_u:.2f} ms", file=sys.stderr, flush=True)
    res = fus.extract(); t_e = fus.timings()
    print(f"[{time.strftime('%X')}] {name}: extract {t_e}", file=sys.stderr, flush=True)
    line = {"config": name, "frames": n_frames, "points_per_frame": npf, "res": g.res, "box": g.box[1] - g.box[0], "dims": list(fus.dims),
            "update_every": update_every, "batch_frames_per_launch": batch,
            "ingest_points_per_s": n_frames * npf / (ingest_ms * 1e-3), "ingest_ms": ingest_ms, "kept_fraction": kept_pts / (n_frames * npf),
            "interleaved_update_ms": upd_ms, "process_ms": t_u + t_e["extract_device_ms"] + t_e["extract_d2h_ms"],
            "update_ms": t_u, "extract_device_ms": t_e["extract_device_ms"], "extract_d2h_ms": t_e["extract_d2h_ms"],
            "occupied_voxels": fus.stats()["occupied_voxels"], "extracted_voxels": len(res)}
    line.update(checks(fus, res, kept_pts, canonical=update_every == 0))
    if og is not None:
        from helpers import RESULT_FIELDS, bits_equal
        t0 = time.perf_counter(); og.update(); want = og.download(); t_proc = time.perf_counter() - t0
        line["oracle_bit_exact"] = all(bits_equal(getattr(res, f), getattr(want, f)) for f in RESULT_FIELDS)
        line["oracle_kind"] = _ORACLE_KIND
        line["oracle_voxels"] = len(want)
        line["result_checksums"] = bench.checksums(res)
        line["oracle_checksums"] = bench.checksums(want)
        line["cpu_points_per_s"] = n_frames * npf / t_cpu
        line["cpu_process_ms"] = t_proc * 1e3
        og.close()
    fus.clear()
    line["clear_empties"] = len(fus.extract()) == 0
    line["data"] = "numpy generator (bit-identical to the oracle's input)" if (oracle or numpy_frames) else "torch generator on the GPU"
    return line


def run_c5(n_sheets, n_side, oracle=False):
    """Extraction stress: stacked one-voxel-thick wavy sheets inserted with pcf_add_points (world frame, explicit viewpoint)."""
    g, sheets = synth.wavy_sheets_world(n_sheets=n_sheets, n_side=n_side)
    fus = pcf.Fusion(g.box, g.res, max_frames=max(1 << 16, n_sheets + 1), log_capacity_hint=sum(len(p) for p, _ in sheets))
    for warm in (True, False):           # first pass sizes every scratch buffer, the second is reported
        t0 = time.perf_counter()
        for i, (pts, vp) in enumerate(sheets):
            fus.add_points(pts, vp, i)
        fus.sync(); t_in = time.perf_counter() - t0
        kept = fus.count_kept()
        fus.update(); t_u = fus.timings()["update_ms"]
        if warm:
            fus.extract_raw(); fus.clear()
            continue
        res = fus.extract(); t_e = fus.timings()
    line = {"config": f"C5 extract stress ({n_sheets} sheets x {n_side}^2)", "points": int(sum(len(p) for p, _ in sheets)), "kept": kept,
            "dims": list(fus.dims), "insert_wall_s": t_in, "process_ms": t_u + t_e["extract_device_ms"] + t_e["extract_d2h_ms"], "update_ms": t_u,
            "extract_device_ms": t_e["extract_device_ms"], "extract_d2h_ms": t_e["extract_d2h_ms"],
            "occupied_voxels": fus.stats()["occupied_voxels"], "extracted_voxels": len(res)}
    line.update(checks(fus, res, kept))
    t_w = fus.timings()             # the idempotence check extracted a second time: warm buffers
    line["extract_device_ms_warm"] = t_w["extract_device_ms"]
    line["process_ms_warm_extract"] = t_u + t_w["extract_device_ms"] + t_w["extract_d2h_ms"]
    if oracle:
        import oracle as O
        from helpers import RESULT_FIELDS, bits_equal
        og = O.OracleGrid(g.box, g.res)
        for pts, vp in sheets:
            og.add_points_world(pts, vp)
        t0 = time.perf_counter(); og.update(); want = og.download(); line["cpu_process_ms"] = (time.perf_counter() - t0) * 1e3
        line["oracle_bit_exact"] = all(bits_equal(getattr(res, f), getattr(want, f)) for f in RESULT_FIELDS)
        og.close()
    fus.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["C1", "C2"])
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--oracle-kind", default="oracle", choices=["oracle", "ref_ordered", "ref"])
    ap.add_argument("--c3-frames", type=int, default=1000)
    ap.add_argument("--c5-sheets", type=int, default=100)
    ap.add_argument("--c5-side", type=int, default=1000)
    a = ap.parse_args()
    global _ORACLE_KIND
    _ORACLE_KIND = a.oracle_kind
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for c in a.configs:
        if c == "C1":
            line = run_scene("C1 replay20", synth.sphere_turntable(20), 20, 20, oracle=a.oracle)
        elif c == "C2":
            line = run_scene("C2 turntable200", synth.sphere_turntable(200, rings=2), 200, 200, oracle=a.oracle)
        elif c == "C3":
            line = run_scene(f"C3 sweep{a.c3_frames}", synth.plate_sweep(a.c3_frames), a.c3_frames, 250, oracle=a.oracle, frames_per_gen=250,
                             passes=1 if a.oracle else 2)
        elif c == "C4":
            line = run_scene("C4 hires50", synth.hires_sphere(50), 50, 10, update_every=10, oracle=a.oracle, frames_per_gen=10,
                             passes=1 if a.oracle else 2)
        elif c.startswith("C3x"):       # C3x200: C3 with the fusion frame rotated so that the plate lies in the y-z plane
            nfr = int(c[3:])            # (layout experiment: the same points land in few x-planes of the x-major grid)
            global _PREMUL
            _PREMUL = np.array([[0, 0, 1, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float64)
            line = run_scene(f"C3 sweep rotated ({nfr} frames)", synth.plate_sweep(1000), nfr, 250, oracle=False, frames_per_gen=250)
            _PREMUL = None
        elif c.startswith("C3n"):       # C3n100: the C3 scene cut to 100 frames
            nfr = int(c[3:])
            line = run_scene(f"C3 sweep ({nfr} frames)", synth.plate_sweep(1000), nfr, 250, oracle=False, frames_per_gen=250)
        elif c.startswith("C4n"):       # C4n30: the C4 scene cut to 30 frames (scaling probe of the interleaved path)
            nfr = int(c[3:])
            line = run_scene(f"C4 hires ({nfr} frames)", synth.hires_sphere(50), nfr, 10, update_every=10, oracle=False, frames_per_gen=10, passes=1)
        elif c == "C4small":
            line = run_scene("C4 hires (20 frames)", synth.hires_sphere(50), 20, 10, update_every=10, oracle=False, frames_per_gen=10, passes=1)
        elif c == "C5":
            line = run_c5(a.c5_sheets, a.c5_side, oracle=False)
        elif c == "C5small":
            line = run_c5(4, 200, oracle=a.oracle)
        else:
            raise SystemExit(f"unknown config {c}")
        print(json.dumps(line), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
