"""Split an `ncu --page source --csv` SASS listing of the sweep kernel into phases (at the BAR.SYNCs, the first
FFMA2 and the first STG) and report instructions per warp / stall samples per phase."""
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
phases = [collections.Counter()]; names = ["start"]
seen_ffma2 = seen_stg = False
nw = None
for r in b["rows"]:
    src = r[idx["Source"]]
    n = int(r[idx["Instructions Executed"]]); s = int(r[idx["# Samples"]])
    if nw is None: nw = n
    if "BAR.SYNC" in src and not seen_ffma2:
        phases.append(collections.Counter()); names.append(f"after barrier {len(phases) - 1}")
    elif "FFMA2" in src and not seen_ffma2:
        seen_ffma2 = True; phases.append(collections.Counter()); names.append("descent (first FFMA2 on)")
    elif "STG" in src and seen_ffma2 and not seen_stg:
        seen_stg = True; phases.append(collections.Counter()); names.append("store + stats")
    a = phases[-1]
    a["inst"] += n; a["samples"] += s
    for c in stall_cols: a[c] += int(r[idx[c]])
tot_s = sum(a["samples"] for a in phases)
print(b["name"][:100])
for nm, a in zip(names, phases):
    top = sorted(((c, a[c]) for c in stall_cols), key=lambda x: -x[1])[:4]
    print(f"{nm:28s} inst/warp {a['inst']/nw:8.1f}  samples {a['samples']:6d} ({a['samples']/tot_s*100:4.1f}%)  " +
          ", ".join(f"{c[6:]}={v}" for c, v in top))
