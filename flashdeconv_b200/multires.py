"""Multi-resolution driver: the paper's "resolution horizon" use-case.

Mirrors `aggregate_to_bin_size` / `run_multiscale_analysis` of the reference's
examples/resolution_horizon_analysis.ipynb (cells 7-8) on plain arrays: spots of the finest lattice are summed
into coarser square bins, and every resolution is deconvolved through the device path.  The bin aggregation is a
sparse row-group sum on the host (scipy, as in the notebook); gene counts / coordinates per level are small next to
the finest level, whose deconvolution dominates.
"""
from __future__ import annotations

import time
from typing import Dict, Optional, Sequence

import numpy as np
from scipy import sparse
from scipy.spatial import cKDTree

from .estimator import FlashDeconv


def aggregate_to_bin_size(Y, coords: np.ndarray, target_um: int, original_um: int = 8):
    """Sum spots into square bins of `target_um` (notebook cell 7).  Returns (Y_agg CSR, coords_agg = mean position,
    group = bin index of every input spot).  The pixel spacing is the median nearest-neighbour distance of the first
    10,000 spots, the grid starts at the coordinate minimum, bins are numbered in lexicographic order of their
    "x_y" labels' numeric parts (the notebook sorts the string labels; only the order of the output rows differs)."""
    coords = np.asarray(coords, dtype=np.float64)
    scale = int(target_um) // int(original_um)
    Yc = Y.tocsr() if sparse.issparse(Y) else sparse.csr_matrix(np.asarray(Y))
    if scale <= 1:
        return Yc, coords.copy(), np.arange(coords.shape[0])
    m = min(10000, coords.shape[0])
    spacing = float(np.median(cKDTree(coords[:m]).query(coords[:m], k=2)[0][:, 1]))
    grid = spacing * scale
    lo = coords.min(axis=0)
    gx = ((coords[:, 0] - lo[0]) / grid).astype(np.int64)
    gy = ((coords[:, 1] - lo[1]) / grid).astype(np.int64)
    key = gx * (int(gy.max()) + 1) + gy
    uniq, group = np.unique(key, return_inverse=True)
    n_bins = uniq.size
    agg = sparse.csr_matrix((np.ones(coords.shape[0], dtype=Yc.dtype), (group, np.arange(coords.shape[0]))),
                            shape=(n_bins, coords.shape[0]))
    size = np.bincount(group, minlength=n_bins).astype(np.float64)
    c_agg = np.column_stack([np.bincount(group, weights=coords[:, d], minlength=n_bins) / size
                             for d in range(coords.shape[1])])
    return (agg @ Yc).tocsr(), c_agg, group


def run_multiscale_analysis(Y, X: np.ndarray, coords: np.ndarray, bin_sizes: Sequence[int] = (8, 16, 32, 64, 128),
                            base_um: int = 8, cell_type_names: Optional[Sequence] = None, random_state: int = 0,
                            **model_kwargs) -> Dict[int, dict]:
    """FlashDeconv at several resolutions (notebook cell 8).  Returns {bin_size: {proportions, coords, n_spots,
    runtime, info, lambda}}.  model_kwargs go to FlashDeconv (the notebook uses lambda_spatial=5000)."""
    out: Dict[int, dict] = {}
    for b in bin_sizes:
        Yb, cb, _ = aggregate_to_bin_size(Y, coords, b, base_um)
        model = FlashDeconv(random_state=random_state, **model_kwargs)
        t0 = time.time()
        prop = model.fit_transform(Yb, X, cb, cell_type_names=cell_type_names)
        out[int(b)] = {"proportions": prop, "coords": cb, "n_spots": int(Yb.shape[0]), "runtime": time.time() - t0,
                       "info": model.info_, "lambda": model.lambda_used_, "dominant": model.get_dominant_cell_type()}
    return out
