"""Per-kernel launch table (markdown) from an `ncu --metrics gpu__time_duration.sum --csv` log of one bench.py step."""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    if r[idx["Metric Name"]] != "gpu__time_duration.sum": continue
    name = re.sub(r"\(.*$", "", r[idx["Kernel Name"]]).strip()
    name = name if name.startswith("void at::") else name.replace("void ", "")
    v = float(r[idx["Metric Value"]]) / (1e3 if r[idx["Metric Unit"]] in ("ns", "nsecond") else 1.0)
    tot[name] += v; cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
for k, v in tot.most_common():
    nm = k if k.startswith(("void at::", "fdb::")) else "fdb::" + k
    print(f"| `{nm}` | {cnt[k]} | {v:.1f} | {v / cnt[k]:.1f} | {100 * v / total:.1f}% |")
print(f"\nTotal kernel time in the step: {total / 1e3:.2f} ms over {sum(cnt.values())} launches.")
