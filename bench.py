#!/usr/bin/env python
"""bench.py -- spots/sec of FlashDeconv's hot path (log-CPM+sketch, graph, lambda, BCD, normalise).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C3]

One "step" = one pass of the whole hot path over one synthetic batch (BASELINE.json config,
default C3: 1,000,000 spots x 18,000 genes, ~2 % density, K=30, kNN k=6, d=512, 100 sweeps).
  value        whole-job spots/s with the CSR / coords / tables already resident in HBM
  e2e          same metric through the host-buffer call (pipeline.deconvolve_path): pinned host CSR ->
               H2D -> path -> float64 beta + proportions D2H, all inside the timed region;
               e2e.public_ms is FlashDeconv.fit_transform(scipy CSR, pageable memory), gene selection included
  roofline     the BCD sweep kernel (dominant): algorithmic bytes / measured average launch time
  cpu_baseline the CPU oracle (numpy/scipy + C/OpenMP port of the reference) on a bounded sample, and
  parity       the GPU path on the SAME sample against that oracle run (north-star bars)
  c5           the 10M-spot configuration (C5, K=50) at this GPU count: ms/step, spots/s, sweep roofline
`--impl reference` times the CPU port alone on the box's host cores over the FULL configuration (the reference is
pure Python + numba and /root/reference does not exist on the GPU box, so the pinned port under oracle/ stands
in).  That arm never imports the product package or touches CUDA: numpy generator, oracle gene selection.
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"] or "--impl=reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every host core (BLAS + OpenMP sweep)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)

import argparse
import importlib.util
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SOLVER = dict(d=512, n_hvg=2000, n_markers=50, rho=0.01, max_iter=100, tol=1e-4, k=6, seed=0)
METRIC = "spots/sec (sketch+graph+BCD)"


def load_synth():
    """flashdeconv_b200/synth.py as a stand-alone module (pure numpy): the reference arm must not import the package."""
    spec = importlib.util.spec_from_file_location("fdb_synth_standalone", os.path.join(ROOT, "flashdeconv_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod          # dataclasses resolve the module through sys.modules
    spec.loader.exec_module(mod)
    return mod


def workload_name(name, cfg):
    return (f"{name}: synthetic {cfg['n_spots']}x{cfg['n_genes']} counts, K={cfg['n_types']}, "
            f"{cfg['method']} graph, d=512, 100 sweeps")


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


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores (reference arm, and the cpu_baseline / parity leg of the b200 arm)
# ---------------------------------------------------------------------------------------------------------
def oracle_run(Y, X, coords, gene_idx, leverage, cfg, max_iter=None):
    from oracle import fd_oracle as fo
    tm = {}
    res = fo.run_path(Y, X, coords, gene_idx, leverage, d=SOLVER["d"], method=cfg["method"], k=SOLVER["k"],
                      rho=SOLVER["rho"], max_iter=SOLVER["max_iter"] if max_iter is None else max_iter,
                      tol=SOLVER["tol"], seed=SOLVER["seed"], timings=tm)
    return res, tm


def run_reference(args, cfg):
    """Times the CPU port on the FULL configuration.  No product import, no CUDA."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from scipy import sparse
    from oracle import fd_oracle as fo
    threads = fo.set_native_threads(os.cpu_count() or 1)
    synth = load_synth()
    n, G = cfg["n_spots"], cfg["n_genes"]
    d = synth.make_dataset_sparse(n, G, cfg["n_types"], cfg["depth"], jitter=cfg["jitter"], seed=SOLVER["seed"])
    Y = sparse.csr_matrix((d["data"].astype(np.float64), d["indices"], d["indptr"]), shape=(n, G))
    X, coords = d["X"], d["coords"]
    del d
    gene_idx, leverage = fo.select_genes(Y, X, SOLVER["n_hvg"], SOLVER["n_markers"])     # step 1: outside the metric
    oracle_run(Y[:2000], X, coords[:2000], gene_idx, leverage, cfg, max_iter=3)           # thread pool, page faults
    vals = []
    for _ in range(args.warmup + args.steps):
        _, tm = oracle_run(Y, X, coords, gene_idx, leverage, cfg)
        vals.append(tm)
    vals = vals[args.warmup:]
    sec = float(np.mean([tm["metric_total"] for tm in vals]))
    v = n / sec
    sample = f"full {args.config} configuration ({n} spots), all stages, 100 sweeps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "spots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, cfg), "nnz": int(Y.nnz), "density": Y.nnz / (n * G),
                   "genes_selected": int(len(gene_idx)), "generator": "numpy (synth.make_dataset_sparse), same model and seed "
                   "as the b200 arm's device generator"},
        "cpu_baseline": {"value": v, "unit": "spots/s", "cores": threads, "kind": "port", "sample": sample,
                         "stages_s": vals[-1]},
        "e2e": {"value": v, "unit": "spots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def cpu_sample_and_parity(data, cfg, gene_idx, leverage, n_sample):
    """CPU port on the first n_sample spots (a contiguous band of lattice rows), and the GPU path on the same band
    through the host-buffer call -- compared with the north-star bars."""
    import torch
    from scipy import sparse
    from oracle import fd_oracle as fo
    from flashdeconv_b200 import pipeline
    threads = fo.set_native_threads(os.cpu_count() or 1)
    ip = data["host_indptr"].numpy()[: n_sample + 1].astype(np.int64)
    ix = data["host_indices"].numpy()[: ip[-1]]
    dv = data["host_data"].numpy()[: ip[-1]]
    Y = sparse.csr_matrix((dv.astype(np.float64), ix, ip), shape=(n_sample, cfg["n_genes"]))
    coords = data["host_coords"].numpy()[:n_sample]
    oracle_run(Y[:2000], data["X"], coords[:2000], gene_idx, leverage, cfg, max_iter=3)   # warm, untimed
    ores, tm = oracle_run(Y, data["X"], coords, gene_idx, leverage, cfg)
    Y32 = sparse.csr_matrix((dv, ix, ip), shape=(n_sample, cfg["n_genes"]))
    res = pipeline.deconvolve_path(Y32, data["X"], coords, gene_idx, leverage, sketch_dim=SOLVER["d"],
                                   spatial_method=cfg["method"], k_neighbors=SOLVER["k"], rho_sparsity=SOLVER["rho"],
                                   max_iter=SOLVER["max_iter"], tol=SOLVER["tol"], random_state=SOLVER["seed"])
    torch.cuda.synchronize()
    A, W = res.graph.to_scipy().tocsr(), ores["A"].tocsr()
    A.sort_indices(); W.sort_indices()
    gp, wp = res.proportions, ores["proportions"]
    pear = []
    for k in range(gp.shape[1]):
        a, b = gp[:, k] - gp[:, k].mean(), wp[:, k] - wp[:, k].mean()
        den = float(np.sqrt((a * a).sum() * (b * b).sum()))
        pear.append(float((a * b).sum() / den) if den > 0 else 1.0)
    parity = {"sample_spots": n_sample, "max_abs_prop": float(np.max(np.abs(gp - wp))), "min_pearson": float(min(pear)),
              "knn_equal": bool(np.array_equal(A.indptr, W.indptr) and np.array_equal(A.indices, W.indices)),
              "n_iter_equal": bool(res.info["n_iterations"] == ores["info"]["n_iterations"]),
              "objective_rel_diff": float(abs(res.info["final_objective"] - ores["info"]["final_objective"]) /
                                          max(abs(ores["info"]["final_objective"]), 1e-300)),
              "bars": "max_abs_prop <= 1e-4, min_pearson >= 0.9999, kNN index sets equal"}
    parity["ok"] = bool(parity["max_abs_prop"] <= 1e-4 and parity["min_pearson"] >= 0.9999 and parity["knn_equal"])
    base = {"value": n_sample / tm["metric_total"], "unit": "spots/s", "cores": threads, "kind": "port",
            "sample": f"first {n_sample} spots of the workload (contiguous lattice band), all stages, 100 sweeps",
            "stages_s": tm}
    return base, parity


# ---------------------------------------------------------------------------------------------------------
# b200 arm
# ---------------------------------------------------------------------------------------------------------
def timed_resident(path, run_kw, steps, warmup, barrier, local_rank, distributed):
    """`steps` passes of the resident hot path; returns (ms_per_step, stage_ms, info, lam, launches, clocks)."""
    import torch
    import torch.distributed as dist
    from flashdeconv_b200._native import lib
    for _ in range(warmup):
        path.run_resident(**run_kw)
    barrier()
    launches0 = lib.fdb_launch_count()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        torch.cuda.profiler.start()          # no-op unless run under `ncu --profile-from-start off`
        start.record()
        for _ in range(steps):
            ev = {}
            _, _, info, lam_used = path.run_resident(events=ev, **run_kw)
        end.record()
        barrier()
        torch.cuda.profiler.stop()
    elapsed_ms = start.elapsed_time(end)
    launches = (lib.fdb_launch_count() - launches0) // steps
    stage_ms = {name: a.elapsed_time(b) for name, (a, b) in ev.items()}
    if distributed:
        tt = torch.tensor([elapsed_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    return elapsed_ms / steps, stage_ms, info, lam_used, int(launches), clocks.summary()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--cpu-sample", type=int, default=100_000, help="spots in the CPU-baseline / parity sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs)")
    ap.add_argument("--no-c5", action="store_true", help="skip the 10M-spot sub-measurement")
    args = ap.parse_args()
    synth = load_synth()
    cfg = synth.CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, cfg)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (inputs are generated on the GPU); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    # host side of the end-to-end legs on the GPU's own NUMA node (pinned buffers are first touched there); the CPU legs
    # below run with the full core set again
    from flashdeconv_b200 import _native
    full_affinity = _native.bind_host_to_gpu(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from flashdeconv_b200 import genes, pipeline
    from flashdeconv_b200.synth import make_dataset_device

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def build_path(csr, coords, tables, K):
        if distributed:
            from flashdeconv_b200 import tiling
            return tiling.TiledPath(csr, coords, tables, K)      # spatial tiles + per-sweep halo exchange
        return pipeline.DevicePath(csr, coords, tables, K)

    def prepare(name, c, pinned):
        # every rank generates the same dataset (same seed) and then sketches / solves only its own spatial tile
        data = make_dataset_device(c["n_spots"], c["n_genes"], c["n_types"], c["depth"], jitter=c["jitter"],
                                   seed=SOLVER["seed"], device=f"cuda:{local_rank}", pinned=pinned)
        csr = pipeline.DeviceCSR(data["indptr"], data["indices"], data["data"], (c["n_spots"], c["n_genes"]))
        gene_idx, leverage = genes.select_informative_genes_device(csr, data["X"], SOLVER["n_hvg"], SOLVER["n_markers"])
        tables = pipeline.build_tables(data["X"], gene_idx, leverage, SOLVER["d"], SOLVER["seed"], c["n_genes"])
        return data, csr, gene_idx, leverage, tables

    def run_kw_of(c):
        return dict(method=c["method"], k=SOLVER["k"], lam="auto", rho=SOLVER["rho"], max_iter=SOLVER["max_iter"],
                    tol=SOLVER["tol"])

    def sweep_roofline(c, K, stage_ms, info, own_rows, deg, config_name):
        peak, peak_src = measured_peak_gbs()
        n_iter = max(info["n_iterations"], 1)
        sweep_bytes = (12 * K + 4 * deg + 4) * own_rows             # H + beta_in + beta_out + graph
        sweep_ms = stage_ms["solve"] / n_iter
        achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "bcd_sweep_p_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": profiled_traffic(config_name, "bcd_sweep"), "peak_source": peak_src,
                "bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_ms}

    # ================= headline configuration ===========================================================
    data, csr, gene_idx, leverage, tables = prepare(args.config, cfg, pinned=True)
    n, G, K = cfg["n_spots"], cfg["n_genes"], cfg["n_types"]
    nnz = csr.nnz
    path = build_path(csr, data["coords"], tables, K)
    ms_per_step, stage_ms, info, lam_used, launches, clocks = timed_resident(
        path, run_kw_of(cfg), args.steps, warmup, barrier, local_rank, distributed)
    value = n / (ms_per_step * 1e-3)                # the N ranks share ONE n-spot problem (strong scaling)
    deg = path.graph.nnz / n
    halo_rows = int(path.plan.n_halo) if distributed else 0
    own_rows = int(path.plan.n_own) if distributed else n
    tiled_mode = getattr(path, "mode_description", "") if distributed else ""
    roof = sweep_roofline(cfg, K, stage_ms, info, own_rows, deg, args.config)
    sketch_bytes = (8 * nnz + 4 * (n + 1) + 4 * (K + 1) * n) * own_rows / n + 4 * SOLVER["d"] * K
    sketch_gbs = sketch_bytes / (stage_ms["sketch"] * 1e-3) / 1e9
    roof["sketch_kernel"] = {"kernel": "sketch_contract_v5_kernel", "achieved": sketch_gbs, "frac": sketch_gbs / roof["peak"],
                             "bytes_per_launch": sketch_bytes, "ms_per_launch": stage_ms["sketch"],
                             "traffic": profiled_traffic(args.config, "sketch_contract")}

    # ---- end to end through the host-buffer call ---------------------------------------------------------
    host = pipeline.HostCSR(data["host_indptr"], data["host_indices"], data["host_data"], (n, G))
    e2e_kw = dict(sketch_dim=SOLVER["d"], spatial_method=cfg["method"], k_neighbors=SOLVER["k"],
                  rho_sparsity=SOLVER["rho"], max_iter=SOLVER["max_iter"], tol=SOLVER["tol"],
                  random_state=SOLVER["seed"], pinned_out=True)
    if distributed:
        path.close()
    del path
    torch.cuda.empty_cache()

    def e2e_call():
        if distributed:
            from flashdeconv_b200 import tiling
            return tiling.deconvolve_path_tiled(host, data["X"], data["host_coords"], gene_idx, leverage, download="rank0",
                                                **e2e_kw)
        return pipeline.deconvolve_path(host, data["X"], data["host_coords"], gene_idx, leverage, **e2e_kw)

    e2e_times, res = [], None
    for it in range(0 if args.no_e2e else 2 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        res = e2e_call()
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_times[2:])) if e2e_times else float("nan")
    if distributed and e2e_times:
        tt = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    h2d = int(getattr(res, "h2d_bytes", 0)) or int(sum(data["host_" + k].numel() * data["host_" + k].element_size()
                                                      for k in ("indptr", "indices", "data", "coords")))
    d2h = int(getattr(res, "d2h_bytes", 0)) or int(2 * n * K * 8)
    e2e = {"value": n / e2e_s if e2e_times else None, "unit": "spots/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s if e2e_times else None,
           "call": "tiling.deconvolve_path_tiled" if distributed else "pipeline.deconvolve_path",
           "bytes_are": "per rank (upload: 1/N of the rows each; download: rank 0 only)" if distributed else "total"}
    # the estimator call a user makes: scipy CSR in pageable memory, gene selection included (single GPU only)
    if not distributed and not args.no_e2e:
        from scipy import sparse
        from flashdeconv_b200 import FlashDeconv
        # fresh numpy copies: ordinary pageable memory, as a user's matrix would be (the host_* tensors are pinned)
        Ysp = sparse.csr_matrix((data["host_data"].numpy().copy(), data["host_indices"].numpy().copy(),
                                 data["host_indptr"].numpy().copy()), shape=(n, G))
        Xn, cn = data["X"], data["host_coords"].numpy()
        pub = []
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            FlashDeconv(sketch_dim=SOLVER["d"], spatial_method=cfg["method"], k_neighbors=SOLVER["k"],
                        random_state=SOLVER["seed"]).fit_transform(Ysp, Xn, cn)
            torch.cuda.synchronize()
            pub.append(time.perf_counter() - t0)
        e2e["public_ms"] = 1e3 * min(pub)
        e2e["public_call"] = "FlashDeconv.fit_transform(scipy CSR, pageable host memory), gene selection included"
        del Ysp

    out = {
        "metric": METRIC, "value": value, "unit": "spots/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, cfg), "nnz": nnz, "density": nnz / (n * G),
                   "genes_selected": int(len(gene_idx)), "mean_degree": deg, "sweeps": info["n_iterations"],
                   "converged": info["converged"], "l2": "inputs_exceed_l2",
                   "multi_gpu": (f"{world} spatial tiles, halo exchange per sweep (rank 0: {own_rows} own + {halo_rows} "
                                 f"halo rows), {tiled_mode}") if world > 1 else "single"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "stage_ms": stage_ms, "roofline": roof,
        "final_objective": info["final_objective"], "lambda": lam_used,
    }

    # ---- CPU baseline + parity on a bounded sample of the same workload (rank 0, single GPU) ---------------
    if full_affinity is not None:
        out["config"]["host_affinity"] = (f"{len(os.sched_getaffinity(0))} GPU-local cores for the GPU legs (NVML affinity of the "
                                          f"device), all {len(full_affinity)} for the CPU legs")
        os.sched_setaffinity(0, full_affinity)
    if not args.no_cpu_baseline and world == 1:
        base, parity = cpu_sample_and_parity(data, cfg, gene_idx, leverage, min(args.cpu_sample, n))
        out["cpu_baseline"], out["parity"] = base, parity

    # ================= 10M-spot configuration at this GPU count ===========================================
    if not args.no_c5 and args.config != "C5":
        del data, csr, tables, res, host
        torch.cuda.empty_cache()
        c5 = synth.CONFIGS["C5"]
        t0 = time.perf_counter()
        data5, csr5, _, _, tables5 = prepare("C5", c5, pinned=False)
        path5 = build_path(csr5, data5["coords"], tables5, c5["n_types"])
        ms5, st5, info5, _, _, _ = timed_resident(path5, run_kw_of(c5), 3, 1, barrier, local_rank, distributed)
        own5 = int(path5.plan.n_own) if distributed else c5["n_spots"]
        roof5 = sweep_roofline(c5, c5["n_types"], st5, info5, own5, path5.graph.nnz / c5["n_spots"], "C5")
        out["c5"] = {"workload": workload_name("C5", c5), "nnz": csr5.nnz, "steps": 3, "warmup": 1, "ms_per_step": ms5,
                     "value": c5["n_spots"] / (ms5 * 1e-3), "unit": "spots/s", "stage_ms": st5,
                     "sweep_us": 1e3 * roof5["ms_per_launch"], "sweep_frac": roof5["frac"],
                     "final_objective": info5["final_objective"], "setup_s": time.perf_counter() - t0}
        if distributed:
            path5.close()

    if distributed:
        from flashdeconv_b200 import tiling
        tiling.release_communicators()
    if rank == 0:
        print(json.dumps(out))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
