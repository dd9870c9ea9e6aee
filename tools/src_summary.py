"""Summarise `ncu --page source --csv` output: instruction mix, stall reasons, hottest lines."""
import csv, collections, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
b = blocks[which]
hdr = b["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
print(len(blocks), "kernel blocks; using", which, b["name"][:90])
tot = 0; byop = collections.Counter(); samp = collections.Counter(); stall = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
nw = None
for r in b["rows"]:
    parts = r[idx["Source"]].split()
    if not parts: continue
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    n = int(r[idx["Instructions Executed"]])
    if nw is None: nw = n
    tot += n; byop[op] += n; samp[op] += int(r[idx["# Samples"]])
    for c in stall_cols: stall[c] += int(r[idx[c]])
print("total warp-inst", tot, " per warp", round(tot / nw, 1), "(warps", nw, ")")
for op, n in byop.most_common(22):
    print(f"{op:12s} {n:>12d} {n / tot * 100:5.1f}%  per-warp {n / nw:8.1f}  samples {samp[op]}")
ts = sum(stall.values())
print([(k, round(v / ts * 100, 1)) for k, v in stall.most_common(10)])
top = sorted(b["rows"], key=lambda r: -int(r[idx["# Samples"]]))[:25]
for r in top:
    print(r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(10), r[idx["Source"]].strip()[:90])
