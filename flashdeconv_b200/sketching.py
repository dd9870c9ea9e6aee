"""Function-level mirror of the reference's ``flashdeconv/core/sketching.py`` over libfdb200.

Same names and argument meaning: ``build_countsketch_matrix`` (:18-84) is host work
(numpy RandomState -> bit-identical buckets/signs), ``project_to_sketch`` (:160-206)
and ``sketch_data`` (:209-260) run the projection on the GPU and hand back float64
numpy arrays like the reference does.  The production path does not go through
these (it uses the fused kernel, pipeline.DevicePath.stage_sketch).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
from scipy import sparse

from .pipeline import _ptr, _stream, countsketch_table, csr_to_device


def build_countsketch_matrix(n_genes: int, sketch_dim: int, leverage_scores: Optional[np.ndarray] = None,
                             random_state=None) -> sparse.csr_matrix:
    bucket, _, weight = countsketch_table(n_genes, sketch_dim, leverage_scores, random_state)
    return sparse.csr_matrix((weight, bucket, np.arange(n_genes + 1)), shape=(n_genes, sketch_dim))


def _single_entry_tables(Omega):
    Om = sparse.csr_matrix(Omega)
    Om.eliminate_zeros()
    per_row = np.diff(Om.indptr)
    if per_row.size and per_row.max() > 1:
        raise NotImplementedError("the B200 projection kernel handles CountSketch matrices (one entry per "
                                  "gene row); denser sketches (method='rademacher') are out of scope")
    bucket = np.full(Om.shape[0], -1, dtype=np.int32)
    weight = np.zeros(Om.shape[0], dtype=np.float32)
    rows = np.flatnonzero(per_row == 1)
    bucket[rows] = Om.indices
    weight[rows] = Om.data
    return bucket, weight


def project_to_sketch(Y_tilde, X_tilde: np.ndarray, Omega):
    """Y_sketch = Y_tilde @ Omega on the GPU (float32 accumulate), X_sketch on the host (float64)."""
    from ._native import check, lib, require_cuda
    torch = require_cuda()
    d = Omega.shape[1]
    bucket, weight = _single_entry_tables(Omega)
    csr = csr_to_device(Y_tilde)
    n = csr.shape[0]
    dev = csr.indices.device
    out = torch.empty((n, d), dtype=torch.float32, device=dev)
    gb, gw = torch.from_numpy(bucket).to(dev), torch.from_numpy(weight).to(dev)
    check(lib.fdb_sketch_project_csr(_ptr(csr.indptr), int(csr.indptr.dtype == torch.int64), _ptr(csr.indices),
                                     _ptr(csr.data), n, csr.shape[1], _ptr(gb), _ptr(gw), d, _ptr(out),
                                     _stream(torch)), "sketch_project_csr")
    X_sketch = np.asarray(X_tilde @ sparse.csr_matrix(Omega))
    return out.cpu().numpy().astype(np.float64), X_sketch


def sketch_data(Y_tilde, X_tilde: np.ndarray, sketch_dim: int = 512, leverage_scores: Optional[np.ndarray] = None,
                method: str = "countsketch", random_state=None):
    if method == "rademacher":
        raise NotImplementedError("method='rademacher' is never selected by FlashDeconv.fit and is out of scope")
    if method != "countsketch":
        raise ValueError(f"Unknown sketching method: {method}")
    Omega = build_countsketch_matrix(Y_tilde.shape[1], sketch_dim, leverage_scores, random_state)
    Ys, Xs = project_to_sketch(Y_tilde, X_tilde, Omega)
    return Ys, Xs, Omega
