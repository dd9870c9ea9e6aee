"""Split an `ncu --page source --csv` SASS listing into phases at marker opcodes and report
instructions executed / stall samples per phase."""
import csv, sys, collections
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
b = blocks[which]; hdr = b["hdr"]; idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
phase = 0; names = ["step0/1 (loads)", "step2 (gather)", "step3 (CD)", "step4+stats"]
agg = [collections.Counter() for _ in names]
seen_bar = False; seen_ffma2 = False; seen_stg = False
nw = None
for r in b["rows"]:
    src = r[idx["Source"]]
    n = int(r[idx["Instructions Executed"]]); s = int(r[idx["# Samples"]])
    if nw is None: nw = n
    if phase == 0 and "BAR.SYNC" in src: phase = 1
    elif phase == 1 and "FFMA2" in src: phase = 2
    elif phase == 2 and "STG" in src: phase = 3
    agg[phase]["inst"] += n; agg[phase]["samples"] += s
    for c in stall_cols: agg[phase][c] += int(r[idx[c]])
tot_s = sum(a["samples"] for a in agg)
for nm, a in zip(names, agg):
    top = sorted(((c, a[c]) for c in stall_cols), key=lambda x: -x[1])[:4]
    print(f"{nm:18s} inst/warp {a['inst']/nw:8.1f}  samples {a['samples']:6d} ({a['samples']/tot_s*100:4.1f}%)  " +
          ", ".join(f"{c[6:]}={v}" for c, v in top))
