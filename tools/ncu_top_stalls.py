"""Top stalled SASS instructions per kernel from `ncu --page source --csv --print-source sass` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
nth = int(sys.argv[3]) if len(sys.argv) > 3 else 0
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 30
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
        cur["data"].append(r)
sel = [b for b in blocks if want in b["name"]]
b = sel[nth]
hdr = b["hdr"]; ix = {n: i for i, n in enumerate(hdr)}; data = b["data"]
print(b["name"][:100], len(data), "instructions")
tot = sum(int(r[ix["# Samples"]]) for r in data)
st = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tt = {n: sum(int(r[ix[n]] or 0) for r in data) for n in st}
print("samples", tot, " ".join(f"{n[6:]}={100*v/tot:.1f}%" for n, v in sorted(tt.items(), key=lambda x: -x[1])[:9]))
for i, r in sorted(enumerate(data), key=lambda x: -int(x[1][ix["# Samples"]]))[:topn]:
    s = {n: int(r[ix[n]] or 0) for n in st}; m = max(s, key=s.get)
    print(str(i).rjust(5), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), m[6:].ljust(12), r[1].strip()[:90])
