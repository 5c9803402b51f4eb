"""Turn gpurun_out/launches.csv (+ prof_full.ncu-rep) into the tracked summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles"); os.makedirs(out, exist_ok=True)
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "launches.csv"))) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows[1:]]
n = len(data); rep = data[2 * n // 3:]            # third repetition = one full step (warm caches aside, ncu serialises)
agg = collections.OrderedDict()
for k, v in rep:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
lines = [f"# ncu launch list, one step of tools/prof_replay.py (200 frames, 1 mm, 0.5 m box): {len(rep)} launches, {tot/1e3:.1f} us total",
         "# (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES)",
         "kernel,launches,total_us,share_pct"]
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{k},{c},{v/1e3:.1f},{100*v/tot:.1f}")
open(os.path.join(out, f"{tag}_launches_summary.csv"), "w").write("\n".join(lines) + "\n")
with open(os.path.join(out, f"{tag}_launches_step.csv"), "w") as f:
    f.write("kernel,duration_ns\n")
    for k, v in rep: f.write(f"{k},{v:.0f}\n")
print("\n".join(lines))
rep_path = os.path.join(ROOT, "gpurun_out", "prof_full.ncu-rep")
if os.path.exists(rep_path):
    raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines())); h, units = rr[0], rr[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    idx = [h.index(w) for w in want]; kn = h.index("Kernel Name")
    with open(os.path.join(out, f"{tag}_ncu_full_summary.csv"), "w") as f:
        f.write("kernel," + ",".join(f"{w} [{units[i]}]" for w, i in zip(want, idx)) + "\n")
        for r in rr[2:]:
            f.write(r[kn].split("(")[0].replace("void ", "") + "," + ",".join(r[i].replace(",", "") for i in idx) + "\n")
            if "k_ingest" in r[kn]:
                def b(i):
                    v = float(r[i].replace(",", "")); u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
                tr = b(h.index("dram__bytes_read.sum")) + b(h.index("dram__bytes_write.sum"))
                json.dump({"kernel": "k_ingest_bulk<16>", "dram_bytes_per_launch": tr, "source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full, 200-frame launch)"},
                          open(os.path.join(out, "ingest_traffic.json"), "w"))
    print(open(os.path.join(out, f"{tag}_ncu_full_summary.csv")).read())
