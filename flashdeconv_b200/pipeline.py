"""Device-resident hot path: sketch -> graph -> BCD, driven through the C ABI.

This is the host-side orchestration that `FlashDeconv.fit` (estimator.py) and
`bench.py` share.  torch is used only for device memory, streams and (in
tiling.py) torch.distributed; every arithmetic step on spots is a libfdb200 call.

Reference steps covered (core/deconv.py:321-398): gene subsetting + log-CPM +
sketch (fused, kernel 1/2), spatial graph (kernel 3), auto lambda, BCD solve
(kernel 4), objective, proportion normalisation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
from scipy import sparse

from . import _native
from ._native import check, lib

GRAPH_MODES = {"knn": 0, "radius": 1, "grid": 2}
MAX_TYPES = 64                   # FDB_MAX_TYPES in include/fdb200.h: register-resident kernels, fused sketch
MAX_TYPES_WIDE = 1024            # FDB_MAX_TYPES_WIDE: warp-per-spot kernels (csrc/wide.cu), single GPU


def _ptr(t):
    if t is None:
        return C.c_void_p(0)
    if not t.is_contiguous():
        raise ValueError("libfdb200 takes dense row-major buffers; got a non-contiguous tensor")
    return C.c_void_p(t.data_ptr())


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@dataclass
class DeviceCSR:
    """Spot-by-gene counts on the device: indptr int64|int32, indices int32, data float32."""
    indptr: "object"
    indices: "object"
    data: "object"
    shape: tuple

    @property
    def nnz(self) -> int:
        return int(self.indices.numel())


@dataclass
class HostCSR:
    """CSR arrays already laid out for the device (pinned torch CPU tensors): skips scipy normalisation."""
    indptr: "object"
    indices: "object"
    data: "object"
    shape: tuple


_STAGING = {}          # device -> (pinned staging buffers, thread pool, copy stream): reused across uploads


def _upload_staged(arr: np.ndarray, torch_dtype, dev, chunk_bytes=64 << 20):
    """Host ndarray in PAGEABLE memory -> device tensor of `torch_dtype`.  A plain cudaMemcpy from pageable memory runs at
    10-12 GB/s (the driver stages through one bounce buffer on one core); here the array goes chunk by chunk through two
    pinned buffers, each chunk filled by several threads (numpy copies release the GIL; the dtype conversion, if any,
    happens in the same pass) while the previous chunk is on the wire.  Small arrays take the plain path."""
    torch = _native.require_cuda()
    flat = np.ascontiguousarray(arr).reshape(-1)
    np_dtype = np.dtype({torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64,
                         torch.float64: np.float64}[torch_dtype])
    n = flat.shape[0]
    if n * np_dtype.itemsize < (32 << 20):
        return torch.from_numpy(np.ascontiguousarray(flat, dtype=np_dtype)).to(dev, non_blocking=True)
    import os
    from concurrent.futures import ThreadPoolExecutor
    key = str(dev)
    if key not in _STAGING:
        bufs = [torch.empty(chunk_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        workers = max(1, min(int(os.environ.get("FDB_UPLOAD_THREADS", "8")), cores))
        _STAGING[key] = (bufs, ThreadPoolExecutor(max_workers=workers), torch.cuda.Stream(dev), workers)
    bufs, pool, stream, workers = _STAGING[key]
    out = torch.empty(n, dtype=torch_dtype, device=dev)
    per = chunk_bytes // np_dtype.itemsize
    views_np = [b.numpy().view(np_dtype) for b in bufs]
    views_t = [b.view(torch_dtype) for b in bufs]
    events = [None, None]
    stream.wait_stream(torch.cuda.current_stream(dev))
    for ci, lo in enumerate(range(0, n, per)):
        k = ci & 1
        if events[k] is not None:
            events[k].synchronize()                       # the buffer's previous chunk has left the host
        m = min(per, n - lo)
        step = -(-m // workers)
        futs = [pool.submit(np.copyto, views_np[k][a:min(a + step, m)], flat[lo + a:lo + min(a + step, m)], "unsafe")
                for a in range(0, m, step)]
        for f in futs:
            f.result()
        with torch.cuda.stream(stream):
            out[lo:lo + m].copy_(views_t[k][:m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        events[k] = ev
    torch.cuda.current_stream(dev).wait_stream(stream)
    for ev in events:                                     # the staging buffers are reused by the next upload
        if ev is not None:
            ev.synchronize()
    return out


def csr_to_device(Y, device="cuda", non_blocking=True) -> DeviceCSR:
    """Accepts a scipy sparse matrix, a dense ndarray (converted to CSR on the host), a HostCSR or a DeviceCSR."""
    torch = _native.require_cuda()
    if isinstance(Y, DeviceCSR):
        return Y
    if isinstance(Y, HostCSR):
        dev = torch.device(device)
        return DeviceCSR(Y.indptr.to(dev, non_blocking=True), Y.indices.to(dev, non_blocking=True),
                         Y.data.to(dev, non_blocking=True), tuple(Y.shape))
    if not sparse.issparse(Y):
        Y = sparse.csr_matrix(np.asarray(Y))
    Y = Y.tocsr()
    if not Y.has_canonical_format:
        Y = Y.copy()
        Y.sum_duplicates()
    ip = np.ascontiguousarray(Y.indptr)
    if ip.dtype not in (np.int32, np.int64):
        ip = ip.astype(np.int64)
    dev = torch.device(device)
    to = lambda a: torch.from_numpy(a).to(dev, non_blocking=non_blocking)
    # indices / counts of any numeric dtype: converted and uploaded in one threaded, double-buffered pass
    return DeviceCSR(to(ip), _upload_staged(Y.indices, torch.int32, dev), _upload_staged(Y.data, torch.float32, dev),
                     tuple(Y.shape))


@dataclass
class SketchTables:
    """Host-built CountSketch tables (core/sketching.py:48-84) expanded to the FULL gene axis."""
    gene_bucket: np.ndarray      # int32[G], -1 for unselected genes
    gene_weight: np.ndarray      # float32[G]
    bucket: np.ndarray           # int64[G_sel]  (bit-exact with the reference's Omega.indices)
    sign: np.ndarray             # int64[G_sel]
    weight: np.ndarray           # float64[G_sel]
    X_sketch: np.ndarray         # float64 K x d
    gram: np.ndarray             # float64 K x K
    d: int
    linear: bool = False         # preprocess "raw" / "pearson": value = count * per-gene factor (no log-CPM)


def countsketch_table(n_genes: int, d: int, leverage, seed):
    """Bucket / sign / weight per selected gene.  numpy's legacy RandomState is called in the
    reference's order (randint, then choice: core/sketching.py:58-59) so draws are bit-identical."""
    from .utils.random import check_random_state
    rng = check_random_state(seed)
    if leverage is None:
        p = np.full(n_genes, 1.0 / max(n_genes, 1))
    else:
        p = np.asarray(leverage, dtype=np.float64)
        p = p / (p.sum() + 1e-10)
    bucket = rng.randint(0, d, size=n_genes)
    sign = rng.choice([-1, 1], size=n_genes)
    entry = sign * np.clip(np.sqrt(p * n_genes + 1e-10), 0.1, 10.0)
    norms = np.maximum(np.sqrt(np.bincount(bucket, weights=entry * entry, minlength=d)), 1e-10)
    weight = entry * (np.sqrt(n_genes / d) / norms[bucket])
    return bucket.astype(np.int64), sign.astype(np.int64), weight


PEARSON_THETA = 100.0            # core/deconv.py:203


def build_tables(X, gene_idx, leverage, d, seed, n_genes_total, preprocess="log_cpm", y_col_mean=None) -> SketchTables:
    """Everything about the reference side that stays on the host (K x G_sel work).

    preprocess follows FlashDeconv._preprocess_data (core/deconv.py:177-235): "log_cpm" transforms values on the
    device; "raw" and "pearson" scale every gene column by a constant (1, or 1 / sigma_g with sigma_g^2 = mu_g +
    mu_g^2 / theta, mu_g = mean over spots + 1e-6 -- `y_col_mean` holds the means of ALL input genes), which is
    folded into the device-side gene weights."""
    X = np.asarray(X, dtype=np.float64)
    gene_idx = np.asarray(gene_idx, dtype=np.intp)
    bucket, sign, weight = countsketch_table(len(gene_idx), d, leverage, seed)
    Xsel = X[:, gene_idx]
    factor = np.ones(len(gene_idx))
    if preprocess == "log_cpm":
        Xt = np.log1p(Xsel / (Xsel.sum(axis=1, keepdims=True) + 1e-10) * 1e4)  # core/deconv.py:194-195
    elif preprocess == "raw":
        Xt = Xsel                                                               # core/deconv.py:227-229
    elif preprocess == "pearson":                                               # core/deconv.py:199-225
        if y_col_mean is None:
            raise ValueError("preprocess='pearson' needs the per-gene means of Y")
        mu_y = np.asarray(y_col_mean, dtype=np.float64)[gene_idx] + 1e-6
        factor = 1.0 / np.sqrt(mu_y + mu_y ** 2 / PEARSON_THETA)
        mu_x = Xsel.mean(axis=0, keepdims=True) + 1e-6
        Xt = Xsel / np.sqrt(mu_x + mu_x ** 2 / PEARSON_THETA)
    else:
        raise ValueError(f"Unknown preprocess method: {preprocess}. Choose from 'log_cpm', 'pearson', or 'raw'.")
    Xs = np.zeros((X.shape[0], d))
    np.add.at(Xs.T, bucket, (Xt * weight).T)                                   # X~ @ Omega, core/sketching.py:202
    gb = np.full(n_genes_total, -1, dtype=np.int32)
    gw = np.zeros(n_genes_total, dtype=np.float32)
    gb[gene_idx] = bucket
    gw[gene_idx] = weight * factor
    return SketchTables(gb, gw, bucket, sign, weight, Xs, Xs @ Xs.T, d, linear=preprocess != "log_cpm")


@dataclass
class DeviceGraph:
    order: "object"          # int32[N]  tile position -> input index
    rank: "object"           # int32[N]  input index -> tile position
    indptr: "object"         # int32[N+1] (tile order)
    indices: "object"        # int32[nnz] (tile-order positions, ascending)
    nnz: int
    radius: Optional[float] = None

    def to_scipy(self):
        """Adjacency in INPUT order with ascending columns, float64 ones (FlashDeconv.adjacency_)."""
        torch = _native.require_cuda()
        n = int(self.order.numel())
        out_ptr = torch.empty(n + 1, dtype=torch.int32, device=self.order.device)
        out_idx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=self.order.device)
        ws = torch.empty(8 * (n + 1) + (1 << 16), dtype=torch.uint8, device=self.order.device)
        check(lib.fdb_graph_to_input_order(_ptr(self.indptr), _ptr(self.indices), _ptr(self.order), _ptr(self.rank),
                                           n, _ptr(out_ptr), _ptr(out_idx), _ptr(ws), ws.numel(), _stream(torch)),
              "graph_to_input_order")
        ip = out_ptr.cpu().numpy()
        ix = out_idx[: self.nnz].cpu().numpy()
        return sparse.csr_matrix((np.ones(self.nnz, dtype=np.float64), ix, ip), shape=(n, n))


def build_graph(coords_dev, method="knn", k=6, radius=None) -> DeviceGraph:
    """coords_dev: float64 CUDA tensor N x D, D = 1, 2 or 3 (input order; utils/graph.py:15-22)."""
    torch = _native.require_cuda()
    if method not in GRAPH_MODES:
        raise ValueError(f"Unknown method: {method}")
    if method == "radius" and radius is None:
        raise ValueError("radius must be specified for radius method")
    if coords_dev.dim() != 2 or coords_dev.shape[1] == 0:
        raise ValueError("coords must be 2D with at least 1 coordinate dimension, "
                         f"got shape {tuple(coords_dev.shape)}")
    dims = int(coords_dev.shape[1])
    if dims > 3:
        raise ValueError(f"libfdb200 builds spatial graphs from 1, 2 or 3 coordinates per spot, got {dims}")
    n = int(coords_dev.shape[0])
    dev = coords_dev.device
    coords_dev = coords_dev.contiguous().to(torch.float64)
    order = torch.empty(n, dtype=torch.int32, device=dev)
    rank = torch.empty(n, dtype=torch.int32, device=dev)
    indptr = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    k_eff = int(k) if method == "knn" else 1
    ws_bytes = int(lib.fdb_graph_workspace_bytes(n, max(k_eff, 1)))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    cap = max(2 * max(k_eff, 1) * n, 16) if method == "knn" else max(12 * n, 16)
    nnz, rad = C.c_int64(0), C.c_double(0.0)
    for _ in range(3):
        indices = torch.empty(cap, dtype=torch.int32, device=dev)
        rc = lib.fdb_graph_build_nd(_ptr(coords_dev), n, dims, GRAPH_MODES[method], k_eff, float(radius or 0.0),
                                    _ptr(order), _ptr(rank), _ptr(indptr), _ptr(indices), cap,
                                    C.byref(nnz), C.byref(rad), _ptr(ws), ws_bytes, _stream(torch))
        if rc == -3 and nnz.value > cap:          # radius graphs: degree unknown up front
            cap = int(nnz.value)
            continue
        check(rc, "graph_build")
        break
    return DeviceGraph(order, rank, indptr, indices, int(nnz.value), rad.value if method != "knn" else None)


def build_sweep_plan(indptr, indices, n_rows: int, nnz: int, n_types: int):
    """Gather plan for the sweep kernel (per-patch halo rows + 16-bit neighbour codes), built once per graph.
    Returns a uint8 CUDA tensor, or None when this row width needs no plan."""
    torch = _native.require_cuda()
    nbytes = int(lib.fdb_bcd_plan_bytes(int(n_rows), int(nnz), int(n_types)))
    if nbytes == 0:
        return None
    plan = torch.empty(nbytes, dtype=torch.uint8, device=indptr.device)
    check(lib.fdb_bcd_plan_build(_ptr(indptr), _ptr(indices), int(n_rows), int(nnz), int(n_types), _ptr(plan), nbytes,
                                 _stream(torch)), "bcd_plan_build")
    return plan


@dataclass
class SolveResult:
    beta: np.ndarray                   # N x K float64, input order
    proportions: np.ndarray            # N x K float64
    info: dict
    lambda_used: float
    graph: DeviceGraph
    tables: SketchTables
    timings: dict = field(default_factory=dict)
    dominant: Optional[np.ndarray] = None      # N int32: argmax over the K types, computed on the device


class DevicePath:
    """One deconvolution on one GPU.  All big buffers are torch CUDA tensors owned by this object."""

    def __init__(self, csr: DeviceCSR, coords_dev, tables: SketchTables, n_types: int):
        self.torch = _native.require_cuda()
        self.csr, self.coords, self.tables, self.K = csr, coords_dev, tables, int(n_types)
        self.Kp = _native.padded_types(self.K)
        if self.K > MAX_TYPES_WIDE:
            raise ValueError(f"at most {MAX_TYPES_WIDE} cell types are supported; got {self.K}")
        self.wide = self.K > MAX_TYPES         # any-K path: unfused sketch + contraction, warp-per-spot sweeps (wide.cu)
        self.dev = csr.indices.device
        t = self.torch
        self.gene_bucket = t.from_numpy(tables.gene_bucket).to(self.dev)
        self.gene_weight = t.from_numpy(tables.gene_weight).to(self.dev)
        # the kernels move sketch rows as 16-byte vectors: the device-side sketch dimension is padded to a multiple of
        # four with empty buckets (no gene maps to them, X_s^T rows are zero), which changes nothing in H or ||y_s||^2
        self.d_dev = 4 * ((tables.d + 3) // 4)
        xst = np.zeros((self.d_dev, self.Kp), dtype=np.float32)
        xst[: tables.d, : self.K] = tables.X_sketch.T
        self.x_sketch_t = t.from_numpy(xst).to(self.dev)
        self.gram32 = np.ascontiguousarray(tables.gram, dtype=np.float32)
        if self.wide:
            gp = np.zeros((self.Kp, self.Kp), dtype=np.float32)
            gp[: self.K, : self.K] = self.gram32
            self.gram_dev = t.from_numpy(gp).to(self.dev)
        n = csr.shape[0]
        self.h = t.empty((n, self.Kp), dtype=t.float32, device=self.dev)
        self.ysq = t.empty(n, dtype=t.float32, device=self.dev)
        self.beta_a = t.empty((n, self.Kp), dtype=t.float32, device=self.dev)
        self.beta_b = t.empty((n, self.Kp), dtype=t.float32, device=self.dev)
        self.state = t.zeros(16, dtype=t.int32, device=self.dev)
        self.graph: Optional[DeviceGraph] = None

    # ---- stages -----------------------------------------------------------------------
    def stage_graph(self, method="knn", k=6, radius=None):
        self.graph = build_graph(self.coords, method, k, radius)
        return self.graph

    def stage_sketch(self):
        """Fused log-CPM + CountSketch + contraction; rows land in tile order (needs the graph)."""
        c, tb = self.csr, self.tables
        row_map = self.graph.rank if self.graph is not None else None
        if self.wide:
            return self._stage_sketch_wide(row_map)
        fn = lib.fdb_sketch_linear_contract_csr if tb.linear else lib.fdb_sketch_contract_csr
        check(fn(_ptr(c.indptr), int(c.indptr.dtype == self.torch.int64), _ptr(c.indices), _ptr(c.data), c.shape[0],
                 c.shape[1], _ptr(self.gene_bucket), _ptr(self.gene_weight), self.d_dev, _ptr(self.x_sketch_t), self.K,
                 _ptr(row_map), _ptr(None), int(len(tb.bucket)), _ptr(self.h), _ptr(self.ysq), _stream(self.torch)),
              "sketch_contract_csr")

    def _stage_sketch_wide(self, row_map, rows_per_pass=262_144):
        """K > 64: the materialising sketch (fdb_sketch_logcpm_csr / fdb_sketch_project_csr) and the chunked contraction
        (fdb_contract), a slab of rows at a time so that Y_s never exceeds rows_per_pass x d floats; rows are then placed in
        tile order."""
        t, c, tb = self.torch, self.csr, self.tables
        n = c.shape[0]
        xs = t.zeros((self.K, self.d_dev), dtype=t.float32, device=self.dev)
        xs[:, : tb.d] = t.from_numpy(np.ascontiguousarray(tb.X_sketch, dtype=np.float32)).to(self.dev)
        fn = lib.fdb_sketch_project_csr if tb.linear else lib.fdb_sketch_logcpm_csr
        is64 = int(c.indptr.dtype == t.int64)
        ys = t.empty((min(n, rows_per_pass), self.d_dev), dtype=t.float32, device=self.dev)
        h_in = t.empty((n, self.Kp), dtype=t.float32, device=self.dev) if row_map is not None else self.h
        q_in = t.empty(n, dtype=t.float32, device=self.dev) if row_map is not None else self.ysq
        for lo in range(0, n, rows_per_pass):
            m = min(rows_per_pass, n - lo)
            ys.zero_()
            # a row slab of the CSR: the row pointers keep their absolute offsets into indices / data
            check(fn(_ptr(c.indptr[lo:]), is64, _ptr(c.indices), _ptr(c.data), m, c.shape[1], _ptr(self.gene_bucket),
                     _ptr(self.gene_weight), self.d_dev, _ptr(ys), _stream(t)), "sketch_rows")
            check(lib.fdb_contract(_ptr(ys), _ptr(xs), m, self.d_dev, self.K, _ptr(h_in[lo:]), _ptr(q_in[lo:]), _stream(t)),
                  "contract")
        if row_map is not None:
            idx = row_map.to(t.int64)
            self.h.index_copy_(0, idx, h_in)
            self.ysq.index_copy_(0, idx, q_in)

    def lambda_auto(self, alpha=0.005) -> float:
        """core/spatial.py:181-190 with mean degree = nnz / N."""
        n = self.csr.shape[0]
        mean_deg = self.graph.nnz / n if n else 0.0
        return float(alpha * float(np.mean(np.diag(self.tables.gram))) / max(mean_deg, 1.0))

    def rho_scaled(self, rho: float) -> float:
        return float(rho) * float(np.mean(np.diag(self.tables.gram)))          # core/solver.py:359-360

    def stage_solve(self, lam: float, rho_scaled: float, max_iter: int, tol: float):
        g = self.graph
        if self.wide:
            self.plan = None
            check(lib.fdb_bcd_solve_wide(_ptr(self.h), _ptr(self.gram_dev), _ptr(self.beta_a), _ptr(self.beta_b),
                                         _ptr(g.indptr), _ptr(g.indices), self.csr.shape[0], self.K, float(lam),
                                         float(rho_scaled), int(max_iter), float(tol), _ptr(self.state), _stream(self.torch)),
                  "bcd_solve_wide")
            return
        self.plan = build_sweep_plan(g.indptr, g.indices, self.csr.shape[0], g.nnz, self.K) if max_iter else None
        check(lib.fdb_bcd_solve(_ptr(self.h), self.gram32.ctypes.data_as(C.c_void_p), _ptr(self.beta_a),
                                _ptr(self.beta_b), _ptr(g.indptr), _ptr(g.indices), self.csr.shape[0], self.K,
                                float(lam), float(rho_scaled), int(max_iter), float(tol), _ptr(self.state),
                                _ptr(self.plan), _stream(self.torch)), "bcd_solve")

    def read_state(self):
        st = self.state.cpu()
        n_iter, conv = int(st[3]), bool(int(st[4]))
        rel = float(st[5:6].view(self.torch.float32)[0])
        return n_iter, conv, rel

    def current_beta(self, n_iter: int):
        return self.beta_a if n_iter % 2 == 0 else self.beta_b

    def objective(self, beta_dev, lam: float, rho_scaled: float) -> float:
        t = self.torch
        g = self.graph
        out = t.zeros(5, dtype=t.float64, device=self.dev)
        if self.wide:
            check(lib.fdb_objective_terms_wide(_ptr(beta_dev), _ptr(self.h), _ptr(self.ysq), _ptr(self.gram_dev),
                                               _ptr(g.indptr), _ptr(g.indices), self.csr.shape[0], self.K, _ptr(out),
                                               _stream(t)), "objective_terms_wide")
        else:
            check(lib.fdb_objective_terms(_ptr(beta_dev), _ptr(self.h), _ptr(self.ysq),
                                          self.gram32.ctypes.data_as(C.c_void_p), _ptr(g.indptr), _ptr(g.indices),
                                          self.csr.shape[0], self.K, _ptr(out), _stream(t)), "objective_terms")
        cross, quad, lap, l1, yty = out.cpu().tolist()
        return 0.5 * (yty - 2.0 * cross + quad) + 0.5 * lam * lap + rho_scaled * l1   # core/solver.py:269-284

    def finish(self, beta_dev, pinned=False):
        t = self.torch
        n = self.csr.shape[0]
        b64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
        p64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
        check(lib.fdb_finish(_ptr(beta_dev), _ptr(self.graph.order), n, self.K, _ptr(b64), _ptr(p64), _stream(t)),
              "finish")
        if pinned:
            hb = t.empty((n, self.K), dtype=t.float64, pin_memory=True)
            hp = t.empty((n, self.K), dtype=t.float64, pin_memory=True)
            hb.copy_(b64, non_blocking=True)
            hp.copy_(p64, non_blocking=True)
            t.cuda.current_stream().synchronize()
            return hb.numpy(), hp.numpy()
        return b64.cpu().numpy(), p64.cpu().numpy()

    # ---- whole path -------------------------------------------------------------------
    def run_resident(self, *, method="knn", k=6, radius=None, lam="auto", rho=0.01, max_iter=100, tol=1e-4,
                     events=None):
        """Hot path with inputs and outputs resident in HBM (what bench.py's `value` times): graph, fused
        sketch, BCD sweeps, objective, float64 beta / proportions in input order.  Returns device tensors.
        ``events``: optional dict that receives (start, end) CUDA-event pairs per stage."""
        t = self.torch
        n = self.csr.shape[0]

        def mark(name, fn):
            if events is None:
                return fn()
            a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            events[name] = (a, b)
            return out

        mark("graph", lambda: self.stage_graph(method, k, radius))
        mark("sketch", self.stage_sketch)
        lam_used = self.lambda_auto() if (isinstance(lam, str) and lam == "auto") else float(lam)
        rho_s = self.rho_scaled(rho)
        mark("solve", lambda: self.stage_solve(lam_used, rho_s, max_iter, tol))
        n_iter, conv, rel = self.read_state()
        beta_dev = self.current_beta(n_iter)
        obj = mark("objective", lambda: self.objective(beta_dev, lam_used, rho_s)) if n else 0.0
        if not hasattr(self, "_b64"):
            self._b64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
            self._p64 = t.empty((n, self.K), dtype=t.float64, device=self.dev)
        mark("finish", lambda: check(lib.fdb_finish(_ptr(beta_dev), _ptr(self.graph.order), n, self.K,
                                                    _ptr(self._b64), _ptr(self._p64), _stream(t)), "finish"))
        info = dict(converged=conv, n_iterations=n_iter, final_objective=obj, objectives=[],
                    final_change=rel if max_iter else 0.0)
        return self._b64, self._p64, info, lam_used

    def run(self, *, method="knn", k=6, radius=None, lam="auto", rho=0.01, max_iter=100, tol=1e-4,
            verbose=False, pinned_out=True) -> SolveResult:
        t = self.torch
        n = self.csr.shape[0]
        self.stage_graph(method, k, radius)
        self.stage_sketch()
        lam_used = self.lambda_auto() if (isinstance(lam, str) and lam == "auto") else float(lam)   # core/deconv.py:370-377
        rho_s = self.rho_scaled(rho)
        objectives = []
        if verbose:
            n_iter, conv, rel = self._solve_verbose(lam_used, rho_s, max_iter, tol, objectives)
        else:
            self.stage_solve(lam_used, rho_s, max_iter, tol)
            n_iter, conv, rel = self.read_state()
        if max_iter == 0:
            rel = 0.0
        beta_dev = self.current_beta(n_iter)
        obj = self.objective(beta_dev, lam_used, rho_s) if n else 0.0
        dom = t.empty(max(n, 1), dtype=t.int32, device=self.dev)
        if n:
            check(lib.fdb_dominant_type(_ptr(beta_dev), _ptr(self.graph.order), n, self.K, _ptr(dom), _stream(t)),
                  "dominant_type")
        beta, prop = self.finish(beta_dev, pinned=pinned_out)
        info = dict(converged=conv, n_iterations=n_iter, final_objective=obj,
                    objectives=objectives if verbose else [], final_change=rel)
        return SolveResult(beta, prop, info, lam_used, self.graph, self.tables, dominant=dom[:n].cpu().numpy())

    def _solve_verbose(self, lam, rho_s, max_iter, tol, objectives):
        """Sweep-at-a-time loop used only for verbose=True (objective every 10 sweeps, core/solver.py:399-404)."""
        g, n = self.graph, self.csr.shape[0]
        st = _stream(self.torch)
        plan = None if self.wide else build_sweep_plan(g.indptr, g.indices, n, g.nnz, self.K)
        check(lib.fdb_bcd_init(_ptr(self.beta_a), n, self.K, _ptr(self.state), st), "bcd_init")
        cur, nxt = self.beta_a, self.beta_b
        n_iter, conv, rel = 0, False, 0.0
        for it in range(max_iter):
            if self.wide:
                check(lib.fdb_bcd_sweep_wide(_ptr(self.h), _ptr(self.gram_dev), _ptr(cur), _ptr(nxt), _ptr(g.indptr),
                                             _ptr(g.indices), n, self.K, float(lam), float(rho_s), float(tol), 1,
                                             _ptr(self.state), st), "bcd_sweep_wide")
            else:
                check(lib.fdb_bcd_sweep(_ptr(self.h), self.gram32.ctypes.data_as(C.c_void_p), _ptr(cur), _ptr(nxt),
                                        _ptr(g.indptr), _ptr(g.indices), n, self.K, float(lam), float(rho_s),
                                        float(tol), 1, _ptr(self.state), _ptr(plan), st), "bcd_sweep")
            n_iter, conv, rel = self.read_state()
            if it % 10 == 0 or it == max_iter - 1:
                obj = self.objective(nxt, lam, rho_s)
                objectives.append(obj)
                print(f"Iteration {it}: objective = {obj:.6f}, rel_change = {rel:.6e}")
            cur, nxt = nxt, cur
            if conv:
                print(f"Converged at iteration {it}")
                break
        return n_iter, conv, rel


def gene_moments(csr: DeviceCSR):
    """Per-gene sum and sum of squares of log1p(CP10k) over all spots (utils/genes.py:52-83) -> host float64."""
    torch = _native.require_cuda()
    G = csr.shape[1]
    sums = torch.zeros(G, dtype=torch.float64, device=csr.indices.device)
    sq = torch.zeros(G, dtype=torch.float64, device=csr.indices.device)
    check(lib.fdb_gene_moments_csr(_ptr(csr.indptr), int(csr.indptr.dtype == torch.int64), _ptr(csr.indices),
                                   _ptr(csr.data), csr.shape[0], G, _ptr(sums), _ptr(sq), _stream(torch)),
          "gene_moments_csr")
    return sums.cpu().numpy(), sq.cpu().numpy()


def group_means(csr: DeviceCSR, labels: np.ndarray, n_groups: int) -> np.ndarray:
    """Mean expression per group of rows of a cells x genes matrix (io/loader.py:119-136) -> host float64
    (n_groups x G).  labels: int array, one group index per row (negative = skip)."""
    torch = _native.require_cuda()
    G = csr.shape[1]
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    sums = torch.zeros((n_groups, G), dtype=torch.float64, device=csr.indices.device)
    lab_d = torch.from_numpy(lab).to(csr.indices.device)
    check(lib.fdb_group_sums_csr(_ptr(csr.indptr), int(csr.indptr.dtype == torch.int64), _ptr(csr.indices), _ptr(csr.data),
                                 _ptr(lab_d), csr.shape[0], G, n_groups, _ptr(sums), _stream(torch)), "group_sums_csr")
    size = np.bincount(lab[lab >= 0], minlength=n_groups).astype(np.float64)
    return sums.cpu().numpy() / np.maximum(size, 1.0)[:, None]


def gene_col_means(csr: DeviceCSR) -> np.ndarray:
    """Per-gene mean of the raw counts over all spots (Y.mean(axis=0), core/deconv.py:207) -> host float64."""
    torch = _native.require_cuda()
    G = csr.shape[1]
    sums = torch.zeros(G, dtype=torch.float64, device=csr.indices.device)
    nnz = int(csr.indices.numel())
    check(lib.fdb_gene_sums_csr(_ptr(csr.indices), _ptr(csr.data), nnz, G, _ptr(sums), _stream(torch)), "gene_sums_csr")
    return sums.cpu().numpy() / max(csr.shape[0], 1)


def deconvolve_path(Y, X, coords, gene_idx, leverage, *, sketch_dim=512, lambda_spatial="auto", rho_sparsity=0.01,
                    spatial_method="knn", k_neighbors=6, radius=None, max_iter=100, tol=1e-4, random_state=0,
                    verbose=False, pinned_out=True, preprocess="log_cpm") -> SolveResult:
    """Steps 2-6 of FlashDeconv.fit for HOST inputs (scipy CSR / ndarray counts, ndarray coords): uploads,
    runs the device path, downloads float64 beta / proportions in input order.  This is the call
    `FlashDeconv.fit` makes after gene selection and the one bench.py times end to end."""
    torch = _native.require_cuda()
    csr = csr_to_device(Y)
    y_mean = gene_col_means(csr) if preprocess == "pearson" else None
    tables = build_tables(X, gene_idx, leverage, sketch_dim, random_state, Y.shape[1], preprocess, y_mean)
    c = coords if torch.is_tensor(coords) else torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64))
    coords_dev = c.to(csr.indices.device, non_blocking=True)
    path = DevicePath(csr, coords_dev, tables, np.asarray(X).shape[0])
    return path.run(method=spatial_method, k=k_neighbors, radius=radius, lam=lambda_spatial, rho=rho_sparsity,
                    max_iter=max_iter, tol=tol, verbose=verbose, pinned_out=pinned_out)
