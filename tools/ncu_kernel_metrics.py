"""Print the interesting raw metrics of every kernel in an .ncu-rep (run here, no GPU needed)."""
import csv, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h, u = rr[0], rr[1]
kn = h.index("Kernel Name")
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stall = [x for x in h if x.startswith("smsp__average_warps_issue_stalled") and x.endswith("per_issue_active.ratio")]
for r in rr[2:]:
    if pat and pat not in r[kn]:
        continue
    print("==", r[kn][:90])
    for k in keys:
        if k in h:
            i = h.index(k); print(f"  {k} [{u[i]}] = {r[i]}")
    st = sorted(((float(r[h.index(x)].replace(',', '') or 0), x) for x in stall), reverse=True)[:7]
    for v, x in st:
        print(f"  stall {x.split('stalled_')[1].split('_per_')[0]} = {v:.2f}")
