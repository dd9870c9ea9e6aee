"""Turns the ncu captures under gpurun_out/ (tools/gpu_profile.sh) into the tracked summaries under profiles/.

    python tools/make_profiles.py [round_tag=r02] [config=C3]
"""
import csv, io, json, os, subprocess, sys, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
cfg = sys.argv[2] if len(sys.argv) > 2 else "C3"
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{cfg}.ncu-rep")
launches = os.path.join(ROOT, "gpurun_out", f"{tag}_launches_{cfg}.csv")
out = os.path.join(ROOT, "profiles")

def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout

raw_csv = run("ncu", "-i", rep, "--page", "raw", "--csv")
src_csv = run("ncu", "-i", rep, "--page", "source", "--csv")
open("/tmp/_raw.csv", "w").write(raw_csv)
open("/tmp/_src.csv", "w").write(src_csv)

# ---- launch list
table = run(sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), launches)
with open(os.path.join(out, f"{tag}_launches_{cfg}.md"), "w") as f:
    f.write(f"# Round 2 -- ncu launch list, one hot-path step on {cfg}\n\n"
            f"Command: `ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv python bench.py "
            f"--config {cfg} --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c5` (tools/gpu_profile.sh).\n"
            "Cold-cache, serialised per-launch times: compare SHARES with bench.py's `stage_ms`, not absolutes.\n\n" + table)

# ---- raw metrics per kernel
rows = list(csv.reader(io.StringIO(raw_csv)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'derived__memory_l1_wavefronts_shared_excessive', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
traffic, seen = {}, set()
with open(os.path.join(out, f"{tag}_ncu_raw_{cfg}.md"), "w") as f:
    f.write(f"# Round 2 -- `ncu --set full --clock-control none` raw metrics, {cfg}\n\n"
            "Captured by tools/gpu_profile.sh (`--import-source on -k regex:\"sketch_contract|bcd_sweep|knn_kernel|objective_kernel\"`), "
            "first launch of every kernel.  Tensor-pipe metrics are 0 by design: no kernel on this path is a GEMM.\n")
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ', '').strip()
        if name in seen:
            continue
        seen.add(name)
        f.write(f"\n## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
        for w in want:
            if w in idx:
                f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        traffic[name] = float(r[idx['dram__bytes_read.sum']]) * mult[units[idx['dram__bytes_read.sum']]] + \
            float(r[idx['dram__bytes_write.sum']]) * mult[units[idx['dram__bytes_write.sum']]]
json.dump({"config": cfg, "source": f"profiles/{tag}_ncu_raw_{cfg}.md (dram__bytes_read.sum + dram__bytes_write.sum per launch)",
           "dram_bytes_per_launch": traffic}, open(os.path.join(out, f"{tag}_traffic_{cfg}.json"), "w"), indent=1)

# ---- source-level summaries of the two hot kernels
blocks = [l.split('","')[1][:60] for l in src_csv.splitlines() if l.startswith('"Kernel Name"')]
def first(prefix):
    for i, b in enumerate(blocks):
        if prefix in b:
            return i
    return None
for prefix, fname, tool in (("sketch_contract", f"{tag}_sketch_source_{cfg}.txt", "src_summary.py"),
                            ("bcd_sweep_p", f"{tag}_sweep_source_{cfg}.txt", "src_summary.py")):
    i = first(prefix)
    if i is None:
        continue
    text = run(sys.executable, os.path.join(ROOT, "tools", tool), "/tmp/_src.csv", str(i))
    if prefix == "bcd_sweep_p":
        text = run(sys.executable, os.path.join(ROOT, "tools", "phase_summary.py"), "/tmp/_src.csv", str(i)) + text
    open(os.path.join(out, fname), "w").write(text)
print("wrote profiles for", tag, cfg, "kernels:", sorted(traffic))
