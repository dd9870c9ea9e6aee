"""Host-side informative-gene selection and leverage scores.

Mirrors the public functions of the reference's ``flashdeconv/utils/genes.py``
(select_hvg :18-145, select_markers :148-235, compute_leverage_scores :238-290,
select_informative_genes :293-341).  Per the north star the K x G_sel SVD and
the G-sized ranking stay on the host; only the O(nnz) moment pass has a device
form (``pipeline.gene_moments``).  Results must equal the reference's gene
indices exactly, because every later stage is conditioned on them.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy import sparse


def _log_cp10k_moments(Y) -> Tuple[np.ndarray, np.ndarray]:
    """mean and ddof=1 variance per gene of log1p(1e4 * Y / max(rowsum, 1))  (genes.py:52-102)."""
    n, g = Y.shape
    if sparse.issparse(Y):
        Yc = Y.tocsr()
        lib = np.maximum(np.asarray(Yc.sum(axis=1)).ravel(), 1.0)
        Z = sparse.diags(10000.0 / lib) @ Yc
        Z.data = np.log1p(Z.data)
        mean = np.asarray(Z.sum(axis=0)).ravel() / n
        if n < 2:
            return mean, np.zeros(g)
        second = np.bincount(Z.indices, weights=Z.data ** 2, minlength=g) / n
        return mean, np.maximum(n / (n - 1) * (second - mean ** 2), 0)
    Yd = np.asarray(Y)
    Z = np.log1p(Yd / np.maximum(Yd.sum(axis=1, keepdims=True), 1) * 10000)
    return Z.mean(axis=0), (Z.var(axis=0, ddof=1) if n >= 2 else np.zeros(g))


def _rank_hvg(mean, var, n_top, min_mean, max_mean, min_disp):
    """z-score the variance inside 20 mean-expression quantile bins, filter, take the top n_top."""
    g = mean.shape[0]
    score = np.zeros(g)
    expressed = mean[mean > 0]
    if expressed.size >= 2:
        cuts = np.unique(np.percentile(expressed, np.linspace(0, 100, 20 + 1)))
        if cuts.size >= 2:
            bin_of = np.clip(np.digitize(mean, cuts) - 1, 0, cuts.size - 2)
            for b in range(cuts.size - 1):
                members = bin_of == b
                if np.count_nonzero(members) > 1:
                    v = var[members]
                    score[members] = (v - np.mean(v)) / (np.std(v) + 1e-10)
    passing = np.flatnonzero((mean >= min_mean) & (mean <= max_mean) & (score >= min_disp))
    if passing.size < n_top:
        chosen = np.argsort(score)[::-1][:n_top]
    else:
        chosen = passing[np.argsort(score[passing])[::-1][:n_top]]
    return np.sort(chosen)


def select_hvg(Y, n_top: int = 2000, min_mean: float = 0.0125, max_mean: float = 3.0,
               min_disp: float = 0.5) -> np.ndarray:
    mean, var = _log_cp10k_moments(Y)
    return _rank_hvg(mean, var, n_top, min_mean, max_mean, min_disp)


def select_markers(X: np.ndarray, n_markers: int = 50, method: str = "diff"):
    """Per-type marker genes ranked by (max - second max) of row-normalised expression."""
    if method != "diff":
        raise ValueError(f"Unknown method: {method}" if method not in ("ratio", "specificity") else
                         f"marker method '{method}' is outside the accelerated path; use 'diff'")
    if n_markers < 0:
        raise ValueError(f"n_markers must be non-negative, got {n_markers}")
    k, g = X.shape
    if n_markers == 0 or k == 0:
        return np.array([], dtype=np.intp), np.array([], dtype=np.intp)
    frac = X / (X.sum(axis=1, keepdims=True) + 1e-10)
    if k == 1:
        idx = np.arange(min(n_markers, g))
        return idx, np.zeros(idx.size, dtype=np.intp)
    ranked = np.sort(frac, axis=0)[::-1]
    gap = ranked[0] - ranked[1]
    best_type = np.argmax(frac, axis=0)
    picks, owners = [], []
    for t in range(k):
        owned = np.flatnonzero(best_type == t)
        if owned.size:
            mine = owned[np.argsort(gap[owned])[::-1][:n_markers]]
        else:
            mine = np.argsort(frac[t])[::-1][:n_markers]
        picks.extend(mine)
        owners.extend([t] * len(mine))
    return np.unique(picks), np.array(owners)


def compute_leverage_scores(X: np.ndarray, regularization: float = 1e-6) -> np.ndarray:
    centred = X - X.mean(axis=0, keepdims=True)
    try:
        U, s, _ = np.linalg.svd(centred.T, full_matrices=False)
    except np.linalg.LinAlgError:
        v = np.var(X, axis=0)
        return v / (v.sum() + regularization)
    r = min(X.shape[0], X.shape[1], s.size)
    pc_weight = s[:r] ** 2 / (s[:r] ** 2 + regularization)
    lev = np.sum(U[:, :r] ** 2 * pc_weight, axis=1)
    return lev / (lev.sum() + regularization)


def select_informative_genes(Y, X: np.ndarray, n_hvg: int = 2000, n_markers_per_type: int = 50):
    hvg = select_hvg(Y, n_top=n_hvg)
    markers, _ = select_markers(X, n_markers=n_markers_per_type)
    gene_idx = np.union1d(hvg, markers).astype(np.intp)
    if gene_idx.size == 0:
        raise ValueError("No genes selected. Increase n_hvg or n_markers_per_type.")
    return gene_idx, compute_leverage_scores(X[:, gene_idx])


def select_informative_genes_device(csr, X: np.ndarray, n_hvg: int = 2000, n_markers_per_type: int = 50):
    """Same selection as `select_informative_genes`, with the O(nnz) moment pass of select_hvg
    (utils/genes.py:52-83) run on the GPU over a `pipeline.DeviceCSR` (float64 accumulation); the G-sized
    binning / ranking, the marker pick and the K x G_sel SVD stay on the host."""
    from . import pipeline
    n = csr.shape[0]
    sums, sumsq = pipeline.gene_moments(csr)
    mean = sums / n
    if n >= 2:
        var = np.maximum(n / (n - 1) * (sumsq / n - mean ** 2), 0)
    else:
        var = np.zeros_like(mean)
    hvg = _rank_hvg(mean, var, n_hvg, 0.0125, 3.0, 0.5)
    markers, _ = select_markers(X, n_markers=n_markers_per_type)
    gene_idx = np.union1d(hvg, markers).astype(np.intp)
    if gene_idx.size == 0:
        raise ValueError("No genes selected. Increase n_hvg or n_markers_per_type.")
    return gene_idx, compute_leverage_scores(X[:, gene_idx])
