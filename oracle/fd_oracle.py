"""
oracle/fd_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (float64) restatement of the reference's hot path, used only as the
checker for the CUDA path: tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  Nothing under
``flashdeconv_b200/`` imports this module.

Each function cites the upstream lines it restates (paths into the reference
repo).  Third-party arithmetic the reference delegates to is delegated to the
SAME libraries here (they ship in this image): scipy.sparse CSR products,
scipy.spatial.cKDTree, numpy's legacy MT19937 RandomState, LAPACK SVD.  The
numba-JIT solver kernels are restated in plain C/OpenMP (bcd_oracle.c).

Pinning: ``oracle/pin_against_reference.py`` imports the real reference in
the build container, checks every function below against it on seeded inputs
(bit-exact for buckets / signs / kNN index sets / gene indices, <=1e-10
relative for float64 sums) and writes tests/golden/*.npz.  Versions used for
the pin: numpy 2.3.5, scipy 1.18.1, numba 0.65.0, reference v0.1.6.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy as np
from scipy import sparse
from scipy.spatial import cKDTree

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libfd_oracle.so")
_lib = None


def build_native(force: bool = False) -> str:
    """Compile bcd_oracle.c -> oracle/_build/libfd_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "bcd_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _LIB_PATH


def _native():
    global _lib
    if _lib is None:
        build_native()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        lib.fdo_bcd_sweep.argtypes = [dp, dp, dp, dp, i64p, i64p, ctypes.c_int64, ctypes.c_int,
                                      ctypes.c_double, ctypes.c_double, dp, dp]
        lib.fdo_bcd_sweep.restype = None
        lib.fdo_normalize.argtypes = [dp, dp, ctypes.c_int64, ctypes.c_int]
        lib.fdo_normalize.restype = None
        lib.fdo_sketch_csr.argtypes = [i64p, i32p, dp, ctypes.c_int64, i32p, dp, ctypes.c_int, dp]
        lib.fdo_sketch_csr.restype = None
        lib.fdo_num_threads.restype = ctypes.c_int
        lib.fdo_set_num_threads.argtypes = [ctypes.c_int]
        lib.fdo_set_num_threads.restype = None
        _lib = lib
    return _lib


def native_threads() -> int:
    return int(_native().fdo_num_threads())


def set_native_threads(n: int) -> int:
    """Thread count of the C/OpenMP sweep (the numba prange of core/solver.py:149), whatever OMP_NUM_THREADS says."""
    _native().fdo_set_num_threads(int(n))
    return native_threads()


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


# --------------------------------------------------------------------------
# a12: consumed preprocessing (host) -- utils/genes.py
# --------------------------------------------------------------------------
def hvg_statistics(Y) -> Tuple[np.ndarray, np.ndarray]:
    """Per-gene mean and ddof=1 variance of log1p(CP10k) -- utils/genes.py:52-102.

    Library size is over ALL genes, floored at 1 (genes.py:57-58)."""
    N, G = Y.shape
    if sparse.issparse(Y):
        Yc = Y.tocsr()
        lib = np.maximum(np.asarray(Yc.sum(axis=1)).ravel(), 1.0)
        Z = sparse.diags(1e4 / lib) @ Yc
        Z.data = np.log1p(Z.data)
        mean = np.asarray(Z.sum(axis=0)).ravel() / N
        if N >= 2:
            sq = np.bincount(Z.indices, weights=Z.data ** 2, minlength=G) / N
            var = np.maximum(N / (N - 1) * (sq - mean ** 2), 0)
        else:
            var = np.zeros(G)
    else:
        Yd = np.asarray(Y)
        tot = np.maximum(Yd.sum(axis=1, keepdims=True), 1)
        Z = np.log1p(Yd / tot * 10000)
        mean = Z.mean(axis=0)
        var = Z.var(axis=0, ddof=1) if N >= 2 else np.zeros(G)
    return mean, var


def hvg_from_statistics(mean, var, n_top=2000, min_mean=0.0125, max_mean=3.0, min_disp=0.5):
    """Binned z-score of the variance and top-n pick -- utils/genes.py:104-145."""
    G = mean.shape[0]
    z = np.zeros(G)
    pos = mean[mean > 0]
    if pos.size >= 2:
        edges = np.unique(np.percentile(pos, np.linspace(0, 100, 21)))
        if edges.size >= 2:
            which = np.clip(np.digitize(mean, edges) - 1, 0, edges.size - 2)
            for b in range(edges.size - 1):
                sel = which == b
                if sel.sum() > 1:
                    v = var[sel]
                    z[sel] = (v - v.mean()) / (v.std() + 1e-10)
    ok = np.where((mean >= min_mean) & (mean <= max_mean) & (z >= min_disp))[0]
    if ok.size < n_top:
        pick = np.argsort(z)[::-1][:n_top]
    else:
        pick = ok[np.argsort(z[ok])[::-1][:n_top]]
    return np.sort(pick)


def marker_genes(X, n_markers=50):
    """'diff' marker selection -- utils/genes.py:148-235 (method="diff" only)."""
    K, G = X.shape
    if n_markers < 0:
        raise ValueError(f"n_markers must be non-negative, got {n_markers}")
    if n_markers == 0 or K == 0:
        return np.array([], dtype=np.intp)
    Xn = X / (X.sum(axis=1, keepdims=True) + 1e-10)
    if K == 1:
        return np.arange(min(n_markers, G))
    s = np.sort(Xn, axis=0)[::-1]
    spec = s[0] - s[1]
    owner = np.argmax(Xn, axis=0)
    out = []
    for k in range(K):
        mine = np.where(owner == k)[0]
        if mine.size:
            out.extend(mine[np.argsort(spec[mine])[::-1][:n_markers]])
        else:
            out.extend(np.argsort(Xn[k])[::-1][:n_markers])
    return np.unique(out)


def leverage_scores(X, reg=1e-6):
    """Leverage of each gene from the SVD of the centred X^T -- utils/genes.py:238-290."""
    K, G = X.shape
    Xc = X - X.mean(axis=0, keepdims=True)
    U, s, _ = np.linalg.svd(Xc.T, full_matrices=False)
    k = min(K, G, s.size)
    w = s[:k] ** 2 / (s[:k] ** 2 + reg)
    lev = ((U[:, :k] ** 2) * w).sum(axis=1)
    return lev / (lev.sum() + reg)


def select_genes(Y, X, n_hvg=2000, n_markers=50):
    """utils/genes.py:293-341 -> (sorted gene_idx, leverage over the selection)."""
    mean, var = hvg_statistics(Y)
    hv = hvg_from_statistics(mean, var, n_top=n_hvg)
    mk = marker_genes(X, n_markers)
    idx = np.union1d(hv, mk).astype(np.intp)
    if idx.size == 0:
        raise ValueError("No genes selected. Increase n_hvg or n_markers_per_type.")
    return idx, leverage_scores(X[:, idx])


# --------------------------------------------------------------------------
# a1: log-CPM -- core/deconv.py:177-197
# --------------------------------------------------------------------------
def log_cpm(Y_sel, X_sel):
    if sparse.issparse(Y_sel):
        lib = np.array(Y_sel.sum(axis=1)).flatten()      # keeps the input dtype, as the reference does
        lib[lib == 0] = 1.0
        Yt = sparse.diags(1e4 / lib) @ Y_sel
        Yt.data = np.log1p(Yt.data)
    else:
        Yt = np.log1p(Y_sel / (Y_sel.sum(axis=1, keepdims=True) + 1e-10) * 1e4)
    Xt = np.log1p(X_sel / (X_sel.sum(axis=1, keepdims=True) + 1e-10) * 1e4)
    return Yt, Xt


# f2: the linear preprocess branches -- core/deconv.py:199-229
def preprocess(Y_sel, X_sel, method="log_cpm"):
    if method == "log_cpm":
        return log_cpm(Y_sel, X_sel)
    if method == "raw":
        return Y_sel.astype(np.float64, copy=False), X_sel.astype(np.float64, copy=False)
    if method == "pearson":
        theta = 100.0
        if sparse.issparse(Y_sel):
            mu = np.asarray(Y_sel.mean(axis=0)).flatten() + 1e-6
            Yt = Y_sel.multiply(1.0 / np.sqrt(mu + mu ** 2 / theta))
        else:
            mu = Y_sel.mean(axis=0, keepdims=True) + 1e-6
            Yt = Y_sel / np.sqrt(mu + mu ** 2 / theta)
        mx = X_sel.mean(axis=0, keepdims=True) + 1e-6
        return Yt, X_sel / np.sqrt(mx + mx ** 2 / theta)
    raise ValueError(f"Unknown preprocess method: {method}. Choose from 'log_cpm', 'pearson', or 'raw'.")


# --------------------------------------------------------------------------
# a2: CountSketch table -- core/sketching.py:48-84
# --------------------------------------------------------------------------
def countsketch_table(n_genes: int, d: int, leverage: Optional[np.ndarray], seed):
    """Returns (bucket[int64], sign[int64], weight[float64]) with Omega[g, bucket[g]] = weight[g].

    Draw order matters for bit-exact buckets/signs: randint first, then choice
    (sketching.py:58-59)."""
    rng = np.random.mtrand._rand if seed is None else (
        seed if isinstance(seed, np.random.RandomState) else np.random.RandomState(seed))
    lev = np.ones(n_genes) / n_genes if leverage is None else leverage / (np.sum(leverage) + 1e-10)
    bucket = rng.randint(0, d, size=n_genes)
    sign = rng.choice([-1, 1], size=n_genes)
    amp = np.clip(np.sqrt(lev * n_genes + 1e-10), 0.1, 10.0)
    raw = sign * amp
    col_norm = np.sqrt(np.bincount(bucket, weights=raw ** 2, minlength=d))
    col_norm = np.maximum(col_norm, 1e-10)
    weight = raw * (np.sqrt(n_genes / d) / col_norm[bucket])
    return bucket.astype(np.int64), sign.astype(np.int64), weight


def omega_matrix(bucket, weight, d):
    n = bucket.shape[0]
    return sparse.csr_matrix((weight, (np.arange(n), bucket)), shape=(n, d))


# --------------------------------------------------------------------------
# a3: projection -- core/sketching.py:190-206
# --------------------------------------------------------------------------
def project(Y_tilde, X_tilde, Omega):
    Ys = Y_tilde @ Omega
    if sparse.issparse(Ys):
        Ys = Ys.toarray()
    Xs = X_tilde @ Omega
    if sparse.issparse(Xs):
        Xs = Xs.toarray()
    return np.asarray(Ys), np.asarray(Xs)


def sketch_full_csr(Y_csr, gene_idx, bucket, weight, d):
    """Fused restatement of (subset -> log-CPM -> project) on the FULL CSR, in C.

    Equivalent to project(log_cpm(Y[:, gene_idx]), Omega) but streams the input
    once -- the same formulation the CUDA kernel uses."""
    Y_csr = Y_csr.tocsr()
    N, G = Y_csr.shape
    gb = np.full(G, -1, dtype=np.int32)
    gw = np.zeros(G, dtype=np.float64)
    gb[gene_idx] = bucket.astype(np.int32)
    gw[gene_idx] = weight
    out = np.zeros((N, d), dtype=np.float64)
    ip = np.ascontiguousarray(Y_csr.indptr, dtype=np.int64)
    ix = np.ascontiguousarray(Y_csr.indices, dtype=np.int32)
    dv = np.ascontiguousarray(Y_csr.data, dtype=np.float64)
    _native().fdo_sketch_csr(_p(ip, ctypes.c_int64), _p(ix, ctypes.c_int32), _p(dv, ctypes.c_double),
                             N, _p(gb, ctypes.c_int32), _p(gw, ctypes.c_double), d,
                             _p(out, ctypes.c_double))
    return out


# --------------------------------------------------------------------------
# a5: spatial graph -- utils/graph.py
# --------------------------------------------------------------------------
def _check_coords(c):
    if c.ndim != 2 or c.shape[1] == 0:
        raise ValueError("coords must be 2D with at least 1 coordinate dimension, "
                         f"got shape {c.shape}")


def knn_adjacency(coords, k=6):
    """utils/graph.py:47-83: cKDTree.query(k+1), drop index==row, A+A^T, binary."""
    _check_coords(coords)
    n = coords.shape[0]
    kk = min(k, n - 1)
    if kk <= 0:
        return sparse.csr_matrix((n, n), dtype=np.float64)
    _, nn = cKDTree(coords).query(coords, k=kk + 1)
    r = np.repeat(np.arange(n), kk + 1)
    c = nn.ravel()
    keep = r != c
    A = sparse.csr_matrix((np.ones(keep.sum()), (r[keep], c[keep])), shape=(n, n))
    A = A + A.T
    A.data[:] = 1.0
    return A


def knn_directed_bruteforce(coords, k):
    """O(N^2) check used on small N: k nearest OTHER points, ties -> smaller index."""
    n = coords.shape[0]
    kk = min(k, n - 1)
    d2 = ((coords[:, None, :] - coords[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d2, np.inf)
    order = np.lexsort((np.broadcast_to(np.arange(n), (n, n)), d2), axis=1)
    return order[:, :kk]


def radius_adjacency(coords, radius):
    """utils/graph.py:108-133: all unordered pairs with distance <= radius."""
    _check_coords(coords)
    n = coords.shape[0]
    pairs = cKDTree(coords).query_pairs(r=radius, output_type="ndarray")
    if len(pairs) == 0:
        return sparse.csr_matrix((n, n), dtype=np.float64)
    r = np.concatenate([pairs[:, 0], pairs[:, 1]])
    c = np.concatenate([pairs[:, 1], pairs[:, 0]])
    return sparse.csr_matrix((np.ones(r.size), (r, c)), shape=(n, n))


def grid_adjacency(coords):
    """utils/graph.py:157-172: radius = 1.5 * median nearest-neighbour distance."""
    _check_coords(coords)
    n = coords.shape[0]
    if n <= 1:
        return sparse.csr_matrix((n, n), dtype=np.float64)
    dist, _ = cKDTree(coords).query(coords, k=2)
    return radius_adjacency(coords, float(np.median(dist[:, 1])) * 1.5)


def adjacency(coords, method="knn", k=6, radius=None):
    """utils/graph.py:175-212 dispatcher."""
    if method == "knn":
        return knn_adjacency(coords, k)
    if method == "radius":
        if radius is None:
            raise ValueError("radius must be specified for radius method")
        return radius_adjacency(coords, radius)
    if method == "grid":
        return grid_adjacency(coords)
    raise ValueError(f"Unknown method: {method}")


# --------------------------------------------------------------------------
# a6/a7: Laplacian and auto-lambda -- core/spatial.py
# --------------------------------------------------------------------------
def laplacian(A):
    """core/spatial.py:57-73 (unnormalised): L = D - A."""
    deg = np.asarray(A.sum(axis=1)).ravel()
    return (sparse.diags(deg) - A).tocsr()


def auto_lambda(X_s, A, alpha=0.005):
    """core/spatial.py:181-190."""
    g = np.mean(np.einsum("kd,kd->k", X_s, X_s))
    mean_deg = np.mean(np.asarray(A.sum(axis=1)).ravel()) if A.shape[0] else 0.0
    return float(alpha * g / max(mean_deg, 1.0))


# --------------------------------------------------------------------------
# a4/a8/a9/a10/a11: solver -- core/solver.py
# --------------------------------------------------------------------------
def objective(beta, H_spot_major, gram, yty, A, lam, rho_scaled):
    """core/solver.py:269-284, with Tr(b^T L b) expanded over the adjacency."""
    cross = float(np.sum(beta * H_spot_major))
    quad = float(np.sum((beta.T @ beta) * gram))
    deg = np.asarray(A.sum(axis=1)).ravel()
    lap = float(np.sum(beta * (deg[:, None] * beta - A @ beta)))
    return 0.5 * (yty - 2.0 * cross + quad) + 0.5 * lam * lap + rho_scaled * float(np.abs(beta).sum())


def normalize(beta):
    """core/solver.py:445-452."""
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    out = np.empty_like(beta)
    if beta.size:
        _native().fdo_normalize(_p(beta, ctypes.c_double), _p(out, ctypes.c_double),
                                beta.shape[0], beta.shape[1])
    return out


def sweep(H_spot_major, gram, b_prev, b_next, ptr, idx, lam, rho_scaled, diffs, absv):
    n, K = b_prev.shape
    _native().fdo_bcd_sweep(_p(H_spot_major, ctypes.c_double), _p(gram, ctypes.c_double),
                            _p(b_prev, ctypes.c_double), _p(b_next, ctypes.c_double),
                            _p(ptr, ctypes.c_int64), _p(idx, ctypes.c_int64), n, K,
                            float(lam), float(rho_scaled), _p(diffs, ctypes.c_double),
                            _p(absv, ctypes.c_double))


def bcd_solve(Y_s, X_s, A, lam=0.1, rho=0.01, max_iter=100, tol=1e-4, trace=None):
    """core/solver.py:330-428.  Returns (beta N x K float64, info dict).

    ``trace``: optional list that receives a copy of beta after the sweeps
    whose 1-based index is in ``trace_at`` (attribute on the list), test-only."""
    n, d = Y_s.shape
    K = X_s.shape[0]
    if n == 0 or K == 0:
        return np.empty((n, K)), dict(converged=True, n_iterations=0, final_objective=0.0,
                                      objectives=[], final_change=0.0)
    gram = np.ascontiguousarray(X_s @ X_s.T)
    H = np.ascontiguousarray((X_s @ Y_s.T).T)          # spot-major copy of the K x N product
    yty = float(np.sum(Y_s ** 2))
    rho_scaled = rho * float(np.mean(np.diag(gram)))
    Ac = A.tocsr()
    ptr = np.ascontiguousarray(Ac.indptr, dtype=np.int64)
    idx = np.ascontiguousarray(Ac.indices, dtype=np.int64)
    cur = np.full((n, K), 1.0 / K)
    nxt = np.empty_like(cur)
    diffs = np.empty(n)
    absv = np.empty(n)
    done, it, rel = False, -1, 0.0
    for it in range(max_iter):
        sweep(H, gram, cur, nxt, ptr, idx, lam, rho_scaled, diffs, absv)
        rel = float(diffs.max() / (absv.max() + 1e-10))
        cur, nxt = nxt, cur
        if trace is not None:
            trace.append(cur.copy())
        if rel < tol:
            done = True
            break
    info = dict(converged=done, n_iterations=it + 1,
                final_objective=objective(cur, H, gram, yty, Ac, lam, rho_scaled),
                objectives=[], final_change=rel)
    return cur, info


# --------------------------------------------------------------------------
# whole path, stage by stage (FlashDeconv.fit steps 2-6, core/deconv.py:321-398)
# --------------------------------------------------------------------------
def run_path(Y, X, coords, gene_idx, leverage, *, d=512, lam="auto", rho=0.01, method="knn",
             k=6, radius=None, max_iter=100, tol=1e-4, seed=0, timings=None, preprocess_method="log_cpm"):
    import time
    t = time.perf_counter
    t0 = t()
    Ysel = Y[:, gene_idx]
    if sparse.issparse(Ysel):
        Ysel = Ysel.tocsr()
    Xsel = X[:, gene_idx]
    t1 = t()
    Yt, Xt = preprocess(Ysel, Xsel, preprocess_method)
    if sparse.issparse(Yt):
        Yt = Yt.tocsr()
    t2 = t()
    bucket, sign, weight = countsketch_table(len(gene_idx), d, leverage, seed)
    Ys, Xs = project(Yt, Xt, omega_matrix(bucket, weight, d))
    t3 = t()
    A = adjacency(np.asarray(coords), method, k, radius)
    t4 = t()
    lam_used = auto_lambda(Xs, A) if isinstance(lam, str) else float(lam)
    t5 = t()
    beta, info = bcd_solve(Ys, Xs, A, lam_used, rho, max_iter, tol)
    prop = normalize(beta)
    t6 = t()
    if timings is not None:
        timings.update(subset=t1 - t0, log_cpm=t2 - t1, sketch=t3 - t2, graph=t4 - t3,
                       auto_lambda=t5 - t4, bcd=t6 - t5, metric_total=t6 - t1)
    return dict(bucket=bucket, sign=sign, weight=weight, Y_s=Ys, X_s=Xs, A=A, lam=lam_used,
                beta=beta, proportions=prop, info=info)
