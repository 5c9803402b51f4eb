"""profiles/sass_summary.txt: per-kernel SASS instruction histogram of libpcfusion.so (cuobjdump -sass), the evidence for what the
kernels are made of: UBLKCP (cp.async.bulk) / SYNCS (mbarrier) in the bulk ingest kernel, FP64 pipe ops (DADD/DMUL/DFMA), XU ops
(MUFU / F2F / I2F / F2I / FRND), atomics (RED / ATOM), warp collectives (VOTE / SHFL / MATCH), no tensor-core ops anywhere.
Run here (no GPU needed):  python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "high-fidelity-pointcloud-fusion_b200", "libpcfusion.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
GROUPS = [("UBLKCP", r"^UBLKCP"), ("UBLKPF", r"^UBLKPF"), ("SYNCS", r"^SYNCS"), ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("RED", r"^RED"), ("ATOM", r"^ATOM"), ("DADD", r"^DADD"), ("DMUL", r"^DMUL"), ("DFMA", r"^DFMA"), ("MUFU", r"^MUFU"), ("F2F", r"^F2F"),
          ("I2F", r"^I2F"), ("F2I", r"^F2I"), ("FRND", r"^FRND"), ("FFMA", r"^FFMA"), ("FADD", r"^FADD"), ("FMUL", r"^FMUL"), ("VOTE", r"^VOTE"),
          ("SHFL", r"^SHFL"), ("MATCH", r"^MATCH"), ("HMMA/IMMA/UTC*MMA (tensor core)", r"^(HMMA|IMMA|DMMA|QMMA|UTCHMMA|UTCIMMA|UTCQMMA|UTCOMMA)"),
          ("LDL/STL (local)", r"^(LDL|STL)")]
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1)); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for name, pat in GROUPS:
            if re.match(pat, op):
                counts[cur][name] += 1
print("# SASS instruction histogram per kernel of libpcfusion.so (sm_100a), static counts; made by tools/sass_summary.py")
print("# kernel | total | " + " | ".join(n for n, _ in GROUPS))
tensor = 0
for k, c in counts.items():
    short = re.sub(r"\(.*", "", k).replace("void ", "").replace("pcf::", "")
    print(f"{short[:70]:70s} | {c['_total']:5d} | " + " | ".join(f"{c[n]:4d}" for n, _ in GROUPS))
    tensor += c["HMMA/IMMA/UTC*MMA (tensor core)"]
print(f"# kernels: {len(counts)}; tensor-core instructions in the whole library: {tensor}")
