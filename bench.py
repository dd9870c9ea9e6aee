#!/usr/bin/env python
"""bench.py -- spots/sec of FlashDeconv's hot path (log-CPM+sketch, graph, lambda, BCD, normalise).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C3]

One "step" = one pass of the whole hot path over one synthetic batch (BASELINE.json config,
default C3: 1,000,000 spots x 18,000 genes, ~2 % density, K=30, kNN k=6, d=512, 100 sweeps).
  value     whole-job spots/s with the CSR / coords / tables already resident in HBM
  e2e       same metric through the host-buffer call (pipeline.deconvolve_path): pinned host CSR ->
            H2D -> path -> float64 beta + proportions D2H, all inside the timed region
  roofline  the BCD sweep kernel (dominant): algorithmic bytes / measured average launch time
  cpu_baseline  the CPU oracle (numpy/scipy + C/OpenMP port of the reference) on a bounded sample
`--impl reference` times that CPU port alone on the box's host cores (the reference is pure Python +
numba; /root/reference does not exist on the GPU box, so the pinned port under oracle/ stands in).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from flashdeconv_b200.synth import CONFIGS  # noqa: E402

SOLVER = dict(d=512, n_hvg=2000, n_markers=50, rho=0.01, max_iter=100, tol=1e-4, k=6, seed=0)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(config, kernel_prefix):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/), or None."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    try:
        for name in sorted(os.listdir(pdir)):
            if name.endswith(f"_traffic_{config}.json"):
                with open(os.path.join(pdir, name)) as f:
                    for k, v in json.load(f)["dram_bytes_per_launch"].items():
                        if k.startswith(kernel_prefix):
                            best = float(v)
    except Exception:
        pass
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def select_genes_device(csr, X, n_spots):
    """Step 1 (not part of the metric): HVG moments on the device, ranking + SVD on the host."""
    from flashdeconv_b200 import genes
    return genes.select_informative_genes_device(csr, X, SOLVER["n_hvg"], SOLVER["n_markers"])


def cpu_sample(data, cfg, gene_idx, leverage, n_sample):
    """Times the CPU port on the first n_sample spots (a contiguous band of lattice rows)."""
    from scipy import sparse
    from oracle import fd_oracle as fo
    ip = data["host_indptr"].numpy()[: n_sample + 1].astype(np.int64)
    ix = data["host_indices"].numpy()[: ip[-1]]
    dv = data["host_data"].numpy()[: ip[-1]].astype(np.float64)
    Y = sparse.csr_matrix((dv, ix, ip), shape=(n_sample, cfg["n_genes"]))
    coords = data["host_coords"].numpy()[:n_sample]
    # warm (thread pool, page faults) on a sliver, untimed
    fo.run_path(Y[:2000], data["X"], coords[:2000], gene_idx, leverage, d=SOLVER["d"], method=cfg["method"],
                k=SOLVER["k"], max_iter=3, seed=SOLVER["seed"])
    tm = {}
    res = fo.run_path(Y, data["X"], coords, gene_idx, leverage, d=SOLVER["d"], method=cfg["method"], k=SOLVER["k"],
                      rho=SOLVER["rho"], max_iter=SOLVER["max_iter"], tol=SOLVER["tol"], seed=SOLVER["seed"],
                      timings=tm)
    return n_sample / tm["metric_total"], tm, res, fo.native_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="spots in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    import torch
    import torch.distributed as dist
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (inputs are generated on the GPU); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    distributed = world > 1 and args.impl == "b200"
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from flashdeconv_b200 import pipeline
    from flashdeconv_b200._native import lib
    from flashdeconv_b200.synth import make_dataset_device

    workload = (f"{args.config}: synthetic {cfg['n_spots']}x{cfg['n_genes']} counts, K={cfg['n_types']}, "
                f"{cfg['method']} graph, d=512, 100 sweeps")
    # every rank generates the same dataset (same seed) and then sketches / solves only its own spatial tile
    data = make_dataset_device(cfg["n_spots"], cfg["n_genes"], cfg["n_types"], cfg["depth"], jitter=cfg["jitter"],
                               seed=SOLVER["seed"], device=f"cuda:{local_rank}", pinned=True)
    n, G, K = cfg["n_spots"], cfg["n_genes"], cfg["n_types"]
    csr = pipeline.DeviceCSR(data["indptr"], data["indices"], data["data"], (n, G))
    nnz = csr.nnz
    gene_idx, leverage = select_genes_device(csr, data["X"], n)

    if args.impl == "reference":
        ns = min(args.cpu_sample, n)
        vals = []
        for _ in range(args.warmup + args.steps):
            v, tm, _, threads = cpu_sample(data, cfg, gene_idx, leverage, ns)
            vals.append((v, tm))
        vals = vals[args.warmup:]
        v = float(np.mean([x[0] for x in vals]))
        sample = f"first {ns} spots of {args.config} (contiguous lattice band), all stages, 100 sweeps"
        print(json.dumps({
            "impl": "reference", "metric": "spots/sec (sketch+graph+BCD)", "value": v, "unit": "spots/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * ns / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "spots/s", "cores": threads, "kind": "port", "sample": sample,
                             "stages_s": vals[-1][1]},
            "e2e": {"value": v, "unit": "spots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    tables = pipeline.build_tables(data["X"], gene_idx, leverage, SOLVER["d"], SOLVER["seed"], G)
    if distributed:
        from flashdeconv_b200 import tiling
        path = tiling.TiledPath(csr, data["coords"], tables, K)      # spatial tiles + per-sweep halo exchange
    else:
        path = pipeline.DevicePath(csr, data["coords"], tables, K)
    run_kw = dict(method=cfg["method"], k=SOLVER["k"], lam="auto", rho=SOLVER["rho"], max_iter=SOLVER["max_iter"],
                  tol=SOLVER["tol"])

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        path.run_resident(**run_kw)
    barrier()
    launches0 = lib.fdb_launch_count()
    stage_ms = {}
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        torch.cuda.profiler.start()          # no-op unless run under `ncu --profile-from-start off`
        start.record()
        for _ in range(args.steps):
            ev = {}
            _, prop_dev, info, lam_used = path.run_resident(events=ev, **run_kw)
            stage_events = ev
        end.record()
        barrier()
        torch.cuda.profiler.stop()
    elapsed_ms = start.elapsed_time(end)
    launches = (lib.fdb_launch_count() - launches0) // args.steps
    for name, (a, b) in stage_events.items():
        stage_ms[name] = a.elapsed_time(b)
    if distributed:
        tt = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    ms_per_step = elapsed_ms / args.steps
    value = n / (ms_per_step * 1e-3)                # strong scaling: the N ranks share ONE n-spot problem

    halo_rows = int(path.plan.n_halo) if distributed else 0
    tiled_mode = {"peer": "direct NVLink peer-memory row pushes + flag/max-norm hand-shake kernel (no NCCL on the data path)",
                  "nccl": "ncclSend/ncclRecv + MAX all-reduce issued from the native loop",
                  "torch": "torch.distributed isend/irecv + all_reduce"}[path.mode] if distributed else ""
    own_rows = int(path.plan.n_own) if distributed else n
    # ---- end to end through the host-buffer call ---------------------------------------
    host = pipeline.HostCSR(data["host_indptr"], data["host_indices"], data["host_data"], (n, G))
    e2e_kw = dict(sketch_dim=SOLVER["d"], spatial_method=cfg["method"], k_neighbors=SOLVER["k"],
                  rho_sparsity=SOLVER["rho"], max_iter=SOLVER["max_iter"], tol=SOLVER["tol"],
                  random_state=SOLVER["seed"], pinned_out=True)
    if distributed:
        path.close()
    del path
    torch.cuda.empty_cache()
    e2e_times = []
    for it in range(0 if args.no_e2e else 2 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        if distributed:
            res = tiling.deconvolve_path_tiled(host, data["X"], data["host_coords"], gene_idx, leverage, **e2e_kw)
        else:
            res = pipeline.deconvolve_path(host, data["X"], data["host_coords"], gene_idx, leverage, **e2e_kw)
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    if args.no_e2e:
        if distributed:
            res = tiling.deconvolve_path_tiled(host, data["X"], data["host_coords"], gene_idx, leverage, **e2e_kw)
        else:
            res = pipeline.deconvolve_path(host, data["X"], data["host_coords"], gene_idx, leverage, **e2e_kw)
        e2e_times = [float("nan")] * 3
    e2e_s = float(np.mean(e2e_times[2:]))
    if distributed:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    h2d = int(sum(data["host_" + k].numel() * data["host_" + k].element_size()
                  for k in ("indptr", "indices", "data", "coords")))
    d2h = int(2 * n * K * 8)

    if distributed:
        tiling.release_communicators()
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (BCD sweep) ------------------------------------
    peak, peak_src = measured_peak_gbs()
    n_iter = max(info["n_iterations"], 1)
    deg = res.graph.nnz / n
    sweep_bytes = (12 * K + 4 * deg + 4) * own_rows             # H + beta_in + beta_out + graph
    sweep_ms = stage_ms["solve"] / n_iter
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    sketch_bytes = (8 * nnz + 4 * (n + 1) + 4 * (K + 1) * n) * own_rows / n + 4 * SOLVER["d"] * K
    sketch_gbs = sketch_bytes / (stage_ms["sketch"] * 1e-3) / 1e9
    out = {
        "metric": "spots/sec (sketch+graph+BCD)", "value": value, "unit": "spots/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "nnz": nnz, "density": nnz / (n * G), "genes_selected": int(len(gene_idx)),
                   "mean_degree": deg, "sweeps": info["n_iterations"], "converged": info["converged"],
                   "l2": "inputs_exceed_l2",
                   "multi_gpu": (f"{world} spatial tiles, halo exchange per sweep (rank 0: {own_rows} own + {halo_rows} "
                                 f"halo rows), {tiled_mode}") if world > 1 else "single"},
        "clocks": clocks.summary(),
        "e2e": {"value": n / e2e_s, "unit": "spots/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s},
        "gpu_launches": int(launches),
        "stage_ms": stage_ms,
        "roofline": {"bound": "hbm", "kernel": "bcd_sweep_p_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": profiled_traffic(args.config, "bcd_sweep"),
                     "peak_source": peak_src,
                     "bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_ms,
                     "sketch_kernel": {"achieved": sketch_gbs, "frac": sketch_gbs / peak,
                                       "bytes_per_launch": sketch_bytes, "ms_per_launch": stage_ms["sketch"]}},
        "final_objective": info["final_objective"], "lambda": lam_used,
    }
    if not args.no_cpu_baseline and world == 1:
        ns = min(args.cpu_sample, n)
        v, tm, ores, threads = cpu_sample(data, cfg, gene_idx, leverage, ns)
        out["cpu_baseline"] = {"value": v, "unit": "spots/s", "cores": threads, "kind": "port",
                               "sample": f"first {ns} spots of {args.config} (contiguous lattice band), all stages, "
                                         "100 sweeps", "stages_s": tm}
    print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
