"""gpurun_out/configs.jsonl (tools/run_configs.py) -> profiles/<tag>_configs.jsonl + a markdown table on stdout."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = {}
for l in open(os.path.join(ROOT, "gpurun_out", "configs.jsonl")):
    d = json.loads(l)
    rows[d["config"]] = d            # the last line per config wins
with open(os.path.join(ROOT, "profiles", f"{tag}_configs.jsonl"), "w") as f:
    for d in rows.values():
        f.write(json.dumps(d) + "\n")
print("| config | input points | kept | GPU ingest (G pts/s) | GPU process() ms (update + extract + D2H) | occupied / extracted voxels | CPU oracle pts/s, process ms | parity |")
print("|---|---|---|---|---|---|---|---|")
for name, d in rows.items():
    pts = d.get("frames", 0) * d.get("points_per_frame", 0) or d.get("points", 0)
    ing = f'{d["ingest_points_per_s"] / 1e9:.1f}' if "ingest_points_per_s" in d else "-"
    kept = f'{d["kept_fraction"]:.2f}' if "kept_fraction" in d else "1.00"
    proc = f'{d["process_ms"]:.2f} ({d["update_ms"]:.2f} + {d["extract_device_ms"]:.2f} + {d["extract_d2h_ms"]:.2f})'
    if d.get("interleaved_update_ms"):
        proc += f'; + {d["interleaved_update_ms"]:.1f} ms in {d["frames"] // d["update_every"]} interleaved updates'
    cpu = f'{d["cpu_points_per_s"] / 1e6:.1f} M, {d["cpu_process_ms"]:.0f}' if "cpu_points_per_s" in d else (f'-, {d["cpu_process_ms"]:.0f}' if "cpu_process_ms" in d else "-")
    par = "bit-exact vs oracle" if d.get("oracle_bit_exact") else "properties: " + ", ".join(k for k in ("x_major_sorted", "buffer_sum_eq_kept", "extract_idempotent", "clear_empties") if d.get(k))
    print(f'| {name} | {pts / 1e6:.1f} M | {kept} | {ing} | {proc} | {d["occupied_voxels"]} / {d["extracted_voxels"]} | {cpu} | {par} |')
