"""Reads a bench.py output stream, picks the JSON line, prints the fields used while iterating on kernels."""
import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
lines = [l for l in sys.stdin if l.lstrip().startswith("{")]
if not lines:
    print(tag, "NO JSON LINE"); sys.exit(0)
d = json.loads(lines[-1])
print(tag, "ms_per_step", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["stage_ms"].items()},
      "sweep_us", round(1e3 * d["roofline"]["ms_per_launch"], 2), "sweep_frac", round(d["roofline"]["frac"], 3),
      "sketch_frac", round(d["roofline"]["sketch_kernel"]["frac"], 3))
e = d.get("e2e", {})
if e.get("ms_per_step"):
    print("   e2e_ms", round(e["ms_per_step"], 2), "h2d", e.get("h2d_bytes_per_step"), "d2h", e.get("d2h_bytes_per_step"), "public_ms", e.get("public_ms"))
if "parity" in d:
    print("   parity", d["parity"])
if "c5" in d:
    c = d["c5"]; print("   c5 ms", round(c["ms_per_step"], 2), {k: round(v, 2) for k, v in c["stage_ms"].items()}, "sweep_us", round(c["sweep_us"], 1), "frac", round(c["sweep_frac"], 3))
print("   objective", d.get("final_objective"))
