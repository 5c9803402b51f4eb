"""Collect the round-2 evidence from gpurun_out/ (scratch) into the tracked profiles/ directory.
    python tools/make_profiles_r02.py <prof_tag> <bench_tag> <configs_tag> <ingest_tag>"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
prof, bench_tag, cfg, ing = (sys.argv[1:5] + ["", "", "", ""])[:4]


def first_json_line(path):
    for l in open(path):
        if l.startswith("{"):
            return json.loads(l)
    raise SystemExit(f"no JSON line in {path}")


def copy_json(src, dst):
    if os.path.exists(os.path.join(G, src)):
        json.dump(first_json_line(os.path.join(G, src)), open(os.path.join(P, dst), "w"), indent=1)
        print("wrote", dst)


# ---- bench lines ---------------------------------------------------------------------------------------------------
if bench_tag:
    copy_json(f"bench_{bench_tag}.json", "r02_bench_n1.json")
    copy_json(f"bench_ref_{bench_tag}.json", "r02_bench_reference_arm.json")
for n, tag in [(2, os.environ.get("N2_TAG", "")), (4, os.environ.get("N4_TAG", "")), (8, os.environ.get("N8_TAG", ""))]:
    if tag:
        copy_json(f"bench_n{n}_{tag}.json", f"r02_bench_n{n}.json")
        copy_json(f"c4_n{n}_{tag}.json", f"r02_c4_sharded_n{n}.json")

# ---- configs ---------------------------------------------------------------------------------------------------------
if cfg:
    lines = [json.loads(l) for l in open(os.path.join(G, f"configs_{cfg}.jsonl")) if l.startswith("{")]
    with open(os.path.join(P, "r02_configs.jsonl"), "w") as f:
        for d in lines:
            f.write(json.dumps(d) + "\n")
    ora = []
    for name in (f"configs_oracle_{cfg}.jsonl", f"configs_ref_{cfg}.jsonl"):
        if os.path.exists(os.path.join(G, name)):
            ora += [json.loads(l) for l in open(os.path.join(G, name)) if l.startswith("{")]
    with open(os.path.join(P, "r02_configs_oracle.jsonl"), "w") as f:
        for d in ora:
            f.write(json.dumps(d) + "\n")
    md = ["# BASELINE configurations on one B200 (tools/run_configs.py, second = warm pass; ingest = HBM-resident clouds, CUDA events)",
          "| config | ingest G pts/s | kept | interleaved updates ms | process() ms (update + extract + D2H) | occupied voxels | extracted | checks |", "|---|---|---|---|---|---|---|---|"]
    for d in lines:
        chk = ", ".join(k for k in ("x_major_sorted", "buffer_sum_eq_kept", "extract_idempotent", "clear_empties") if d.get(k))
        md.append(f"| {d['config']} | {d.get('ingest_points_per_s', 0) / 1e9:.0f} | {d.get('kept_fraction', 0):.3f} | {d.get('interleaved_update_ms', 0):.2f} | "
                  f"{d['process_ms']:.2f} ({d['update_ms']:.2f} + {d['extract_device_ms']:.2f} + {d['extract_d2h_ms']:.2f}) | {d['occupied_voxels']} | {d['extracted_voxels']} | {chk} |")
    md += ["", "# Full-size comparisons with the CPU checkers (same numpy-generated frames fed to both; --oracle): every output field compared bitwise",
           "| config | checker | voxels | bit-exact | CPU ingest M pts/s (1 core) | CPU process() ms |", "|---|---|---|---|---|---|"]
    for d in ora:
        md.append(f"| {d['config']} | {d['oracle_kind']} | {d['oracle_voxels']} | {d['oracle_bit_exact']} | {d['cpu_points_per_s'] / 1e6:.1f} | {d['cpu_process_ms']:.0f} |")
    open(os.path.join(P, "r02_configs.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))

# ---- ncu launch list + full capture --------------------------------------------------------------------------------
def ncu_raw(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    return rr[0], rr[1], rr[2:]


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def full_summary(rep, out_csv, traffic_json=None):
    h, units, rows = ncu_raw(rep)
    idx = [h.index(w) for w in WANT if w in h]
    names = [w for w in WANT if w in h]
    kn = h.index("Kernel Name")
    with open(out_csv, "w") as f:
        f.write("kernel," + ",".join(f"{w} [{units[i]}]" for w, i in zip(names, idx)) + "\n")
        for r in rows:
            f.write('"' + r[kn].split("(")[0].replace("void ", "") + '",' + ",".join(r[i].replace(",", "") for i in idx) + "\n")
            if traffic_json and "k_ingest_bulk" in r[kn]:
                def b(name):
                    i = h.index(name); v = float(r[i].replace(",", "")); u = units[i].lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
                json.dump({"kernel": "k_ingest_bulk<16>", "dram_bytes_per_launch": b("dram__bytes_read.sum") + b("dram__bytes_write.sum"),
                           "source": f"profiles/{os.path.basename(out_csv)} (ncu --set full, 200-frame launch)"}, open(traffic_json, "w"))
    print("wrote", out_csv)


if prof:
    rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{prof}.csv"))) if len(r) > 5]
    hi = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    hdr = rows[hi]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if r[vi] not in ("", "Metric Value")]
    starts = [i for i, (k, _) in enumerate(data) if "k_ingest_bulk" in k]
    rep = data[starts[-1]:]                       # the last repetition = one full step
    agg = collections.OrderedDict()
    for k, v in rep:
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# ncu launch list, one step of tools/prof_replay.py (200 frames in one launch, 1 mm, 0.5 m box -> update -> extract -> clear): {len(rep)} launches, {tot / 1e3:.1f} us total",
             "# (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES)",
             "kernel,launches,total_us,share_pct"]
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        lines.append(f'"{k}",{c},{v / 1e3:.1f},{100 * v / tot:.1f}')
    open(os.path.join(P, "r02_launches_summary.csv"), "w").write("\n".join(lines) + "\n")
    with open(os.path.join(P, "r02_launches_step.csv"), "w") as f:
        f.write("kernel,duration_ns\n")
        for k, v in rep:
            f.write(f'"{k}",{v:.0f}\n')
    print("\n".join(lines[:14]))
    full_summary(os.path.join(G, f"prof_full_{prof}.ncu-rep"), os.path.join(P, "r02_ncu_full_summary.csv"), os.path.join(P, "ingest_traffic.json"))

if ing:
    out = ["# ncu --set full rows of the ingest kernel on the C1 (one 20-frame launch, 500^3 grid) and C4 (10-frame launches of 1920x1080 clouds,",
           "# 1000^3 grid at 0.5 mm) shapes; live CUDA-event timings of the same launches next to them (tools/prof_ingest_configs.py)", ""]
    for c in ("C1", "C4"):
        live = [json.loads(l) for l in open(os.path.join(G, f"ingest_{c}_{ing}.jsonl")) if l.startswith("{")]
        out.append(f"## {c}: live launches " + "; ".join(f"{d['ms'] * 1e3:.1f} us = {d['points_per_s'] / 1e9:.0f} G pts/s" for d in live))
        h, units, rows = ncu_raw(os.path.join(G, f"prof_ingest_{c}_{ing}.ncu-rep"))
        for r in rows:
            for w in WANT:
                if w in h:
                    out.append(f"  {w} [{units[h.index(w)]}] = {r[h.index(w)]}")
        out.append("")
    open(os.path.join(P, "r02_ingest_c1_c4.md"), "w").write("\n".join(out) + "\n")
    print("wrote r02_ingest_c1_c4.md")
