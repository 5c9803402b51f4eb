"""profiles/<tag>_ncu_full_summary.csv -> profiles/<tag>_kernel_roofline.md: achieved DRAM GB/s of every captured kernel
against the measured HBM peak (MEASURED_PEAKS.json, 6545.3 GB/s on this pool)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01d"
peak = 6545.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
rows = list(csv.reader(open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv"))))
h = rows[0]
nnum = len(h) - 1


def col(name):
    return [i for i, x in enumerate(h) if x.startswith(name)][0]


def val(r, i):
    u = h[i].split("[")[1].rstrip("]").lower() if "[" in h[i] else ""
    return float(r[i]) * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1}.get(u, 1)


it, ir, iw = col("gpu__time_duration"), col("dram__bytes_read"), col("dram__bytes_write")
io, ig, ih = col("sm__warps_active"), col("launch__registers"), col("lts__t_sector_hit_rate")
out = [f"# Per-kernel DRAM throughput of one C2 step (ncu --set full, profiles/{tag}_ncu_full_summary.csv); peak = measured copy bandwidth {peak:.0f} GB/s",
       "| kernel | time (us) | DRAM read (MB) | DRAM write (MB) | achieved GB/s | of measured peak | warps active % | regs | L2 hit % |",
       "|---|---|---|---|---|---|---|---|---|"]
for r in rows[1:]:
    r = [",".join(r[:len(r) - nnum])] + r[len(r) - nnum:]      # kernel names contain commas
    t, rd, wr = val(r, it), val(r, ir), val(r, iw)
    bw = (rd + wr) / t / 1e9
    out.append(f"| `{r[0]}` | {t * 1e6:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {bw:.0f} | {bw / peak:.2f} | {float(r[io]):.0f} | {r[ig]} | {float(r[ih]):.0f} |")
open(os.path.join(ROOT, "profiles", f"{tag}_kernel_roofline.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
