"""Deterministic synthetic spatial-transcriptomics inputs (spots x genes counts).

Recipe follows the reference's own generators (tests/test_integration.py:44-84,
examples/quickstart.py:24-59) adapted to sparse output: log-normal signatures
with 20 up-regulated marker genes per cell type, a jittered square lattice of
spot coordinates (jitter keeps k-NN tie-free), exponentially decaying spatial
mixing proportions, Gamma-distributed depth and Poisson counts emitted chunk by
chunk straight to CSR.

``make_dataset`` is pure numpy (small / parity cases, identical bits on every
host).  ``make_dataset_device`` draws the counts with torch on a CUDA device
for the million-spot benchmark shapes, where a host generator would take
minutes; it follows the same statistical model.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
from scipy import sparse


@dataclass
class SyntheticData:
    Y: sparse.csr_matrix          # N x G counts, float32 data, int32 indices
    X: np.ndarray                 # K x G signatures, float64
    coords: np.ndarray            # N x 2 float64
    beta_true: np.ndarray         # N x K float64 (rows sum to 1)

    @property
    def density(self) -> float:
        return self.Y.nnz / (self.Y.shape[0] * self.Y.shape[1])


# Named BASELINE.json configurations: (N, G, K, mean counts per spot, lattice jitter, graph method)
CONFIGS = {
    "C1": dict(n_spots=10_000, n_genes=2_000, n_types=10, depth=5000.0, jitter=0.1, method="knn"),
    "C2": dict(n_spots=100_000, n_genes=18_000, n_types=20, depth=400.0, jitter=0.1, method="knn"),
    "C3": dict(n_spots=1_000_000, n_genes=18_000, n_types=30, depth=400.0, jitter=0.1, method="knn"),
    "C4": dict(n_spots=2_000_000, n_genes=25_000, n_types=40, depth=270.0, jitter=0.0, method="grid"),
    "C5": dict(n_spots=10_000_000, n_genes=18_000, n_types=50, depth=92.0, jitter=0.1, method="knn"),
}


def _signatures(rng, n_types, n_genes):
    X = np.exp(rng.standard_normal((n_types, n_genes)) * 0.5 + 1.0)
    for k in range(n_types):
        X[k, rng.choice(n_genes, size=min(20, n_genes), replace=False)] *= 5.0
    return X


def _coords(rng, n_spots, jitter):
    side = int(np.ceil(np.sqrt(n_spots)))
    gx = np.tile(np.arange(side), side)[:n_spots]
    gy = np.repeat(np.arange(side), side)[:n_spots]
    c = np.column_stack([gx, gy]).astype(np.float64)
    if jitter > 0:
        c += rng.standard_normal((n_spots, 2)) * jitter
    return c, side


def _mixing(rng, coords, side, n_types):
    centers = rng.random((n_types, 2)) * side
    b = np.empty((coords.shape[0], n_types))
    for k in range(n_types):
        b[:, k] = np.exp(-np.sqrt(((coords - centers[k]) ** 2).sum(1)) / (side / 2))
    return b / b.sum(1, keepdims=True)


def make_dataset(n_spots=1000, n_genes=500, n_types=5, depth=2000.0, jitter=0.1, seed=0,
                 chunk=20_000, method=None) -> SyntheticData:
    """Host (numpy) generator; deterministic for a given seed."""
    rng = np.random.default_rng(seed)
    X = _signatures(rng, n_types, n_genes)
    coords, side = _coords(rng, n_spots, jitter)
    beta = _mixing(rng, coords, side, n_types)
    spot_depth = rng.gamma(shape=5.0, scale=depth / 5.0, size=n_spots)
    Xp = X / X.sum(1, keepdims=True)
    blocks = []
    for lo in range(0, n_spots, chunk):
        hi = min(lo + chunk, n_spots)
        rate = (beta[lo:hi] @ Xp) * spot_depth[lo:hi, None]
        blocks.append(sparse.csr_matrix(rng.poisson(rate).astype(np.float32)))
    Y = sparse.vstack(blocks, format="csr") if blocks else sparse.csr_matrix((0, n_genes), dtype=np.float32)
    Y.indices = Y.indices.astype(np.int32)
    Y.indptr = Y.indptr.astype(np.int64)
    return SyntheticData(Y=Y, X=X, coords=coords, beta_true=beta)


def make_dataset_sparse(n_spots, n_genes, n_types, depth, jitter=0.1, seed=0, chunk=25_000, threads=None):
    """Host (numpy) generator for LARGE sparse shapes: the same statistical model as `make_dataset`, sampled molecule by
    molecule so that the cost is O(total counts) instead of O(N x G) -- T_i ~ Poisson(depth_i) molecules per spot, each
    with a type ~ beta_i and a gene ~ the type's profile (equivalent to independent Poisson(depth_i * (beta_i Xp)_g)
    counts).  Chunks are drawn by a thread pool from spawned seed sequences: deterministic for a given (seed, chunk)
    whatever the thread count.  Returns a dict of numpy arrays: indptr int64, indices int32, data float32, coords, X,
    beta_true, shape.  1M x 18k at depth 400 takes ~1 minute on 8 cores."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(seed)
    X = _signatures(rng, n_types, n_genes)
    coords, side = _coords(rng, n_spots, jitter)
    beta = _mixing(rng, coords, side, n_types)
    spot_depth = rng.gamma(shape=5.0, scale=depth / 5.0, size=n_spots)
    cdf = np.cumsum(X / X.sum(1, keepdims=True), axis=1)
    cdf /= cdf[:, -1:]
    cdf_flat = (cdf + np.arange(n_types)[:, None]).ravel()              # type k occupies [k, k + 1)
    n_chunks = (n_spots + chunk - 1) // chunk
    seeds = np.random.SeedSequence(seed).spawn(max(n_chunks, 1))

    def one(ci):
        r = np.random.default_rng(seeds[ci])
        lo, hi = ci * chunk, min((ci + 1) * chunk, n_spots)
        T = r.poisson(spot_depth[lo:hi])
        per_type = r.multinomial(T, beta[lo:hi])                        # molecules per (spot, type)
        types = np.repeat(np.tile(np.arange(n_types), hi - lo), per_type.ravel())
        spots = np.repeat(np.arange(hi - lo, dtype=np.int64), T)
        genes = np.searchsorted(cdf_flat, r.random(types.size) + types, side="right") - types * n_genes
        np.clip(genes, 0, n_genes - 1, out=genes)
        key = spots * n_genes + genes
        key.sort()
        first = np.empty(key.size, dtype=bool)
        first[:1] = True
        np.not_equal(key[1:], key[:-1], out=first[1:])
        starts = np.flatnonzero(first)
        cnt = np.diff(np.append(starts, key.size)).astype(np.float32)
        uk = key[starts]
        rows = uk // n_genes
        return np.bincount(rows, minlength=hi - lo), (uk - rows * n_genes).astype(np.int32), cnt

    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as ex:
        parts = list(ex.map(one, range(n_chunks)))
    indptr = np.zeros(n_spots + 1, dtype=np.int64)
    if parts:
        np.cumsum(np.concatenate([p[0] for p in parts]), out=indptr[1:])
    cat = lambda i, dt: np.concatenate([p[i] for p in parts]) if parts else np.zeros(0, dtype=dt)
    return dict(indptr=indptr, indices=cat(1, np.int32), data=cat(2, np.float32), coords=coords, X=X, beta_true=beta,
                shape=(n_spots, n_genes))


def make_dataset_device(n_spots, n_genes, n_types, depth, jitter=0.1, seed=0, chunk=32_768,
                        device="cuda", method=None, pinned=False):
    """CUDA-side generator for benchmark shapes.

    Returns a dict of DEVICE torch tensors (indptr int64, indices int32, data float32,
    coords float64 N x 2) plus host numpy X / beta_true.  With ``pinned=True`` it also
    returns pinned host copies of the CSR arrays and coords under the ``host_*`` keys."""
    import torch

    rng = np.random.default_rng(seed)
    X = _signatures(rng, n_types, n_genes)
    coords, side = _coords(rng, n_spots, jitter)
    beta = _mixing(rng, coords, side, n_types)
    spot_depth = rng.gamma(shape=5.0, scale=depth / 5.0, size=n_spots)
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    Xp = torch.as_tensor(X / X.sum(1, keepdims=True), dtype=torch.float32, device=dev)
    beta_d = torch.as_tensor(beta, dtype=torch.float32, device=dev)
    depth_d = torch.as_tensor(spot_depth, dtype=torch.float32, device=dev)
    counts_per_row, idx_parts, val_parts = [], [], []
    for lo in range(0, n_spots, chunk):
        hi = min(lo + chunk, n_spots)
        rate = (beta_d[lo:hi] @ Xp) * depth_d[lo:hi, None]
        cnt = torch.poisson(rate, generator=gen)
        nz = cnt.nonzero(as_tuple=False)                      # row-major order -> sorted columns per row
        counts_per_row.append(torch.bincount(nz[:, 0], minlength=hi - lo))
        idx_parts.append(nz[:, 1].to(torch.int32))
        val_parts.append(cnt[nz[:, 0], nz[:, 1]])
        del rate, cnt, nz
    row_nnz = torch.cat(counts_per_row)
    indptr = torch.zeros(n_spots + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(row_nnz, 0)
    out = dict(indptr=indptr, indices=torch.cat(idx_parts), data=torch.cat(val_parts),
               coords=torch.as_tensor(coords, device=dev), X=X, beta_true=beta,
               shape=(n_spots, n_genes))
    if pinned:
        for key in ("indptr", "indices", "data", "coords"):
            host = torch.empty(out[key].shape, dtype=out[key].dtype, pin_memory=True)
            host.copy_(out[key])
            out["host_" + key] = host
    return out
