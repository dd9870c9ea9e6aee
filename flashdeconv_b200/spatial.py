"""Host shims for the reference's ``flashdeconv/core/spatial.py``.

Only ``auto_tune_lambda`` (:144-192) and the unnormalised ``compute_laplacian``
(:35-73) are semantically on the hot path, and on the device neither exists as a
matrix: the Laplacian quadratic form is evaluated from the adjacency inside
fdb_objective_terms and lambda needs only nnz/N.  These scipy versions keep the
reference's function surface for callers and tests; they are O(nnz) host code.
"""
from __future__ import annotations

from typing import List

import numpy as np
from scipy import sparse


def compute_degree_matrix(A):
    return sparse.diags(np.asarray(A.sum(axis=1)).ravel(), format="dia")


def compute_laplacian(A, normalized: bool = False) -> sparse.csr_matrix:
    deg = np.asarray(A.sum(axis=1)).ravel()
    if not normalized:
        return (sparse.diags(deg) - A).tocsr()
    scale = np.zeros_like(deg)
    scale[deg > 0] = 1.0 / np.sqrt(deg[deg > 0])
    S = sparse.diags(scale)
    return (sparse.eye(A.shape[0]) - S @ A @ S).tocsr()


def get_neighbor_indices(A) -> List[np.ndarray]:
    Ac = A.tocsr()
    return [Ac.indices[Ac.indptr[i]:Ac.indptr[i + 1]].copy() for i in range(Ac.shape[0])]


def get_neighbor_counts(A) -> np.ndarray:
    return np.asarray(A.sum(axis=1)).ravel().astype(np.int32)


def compute_laplacian_quadratic(beta: np.ndarray, L) -> float:
    return float(np.sum(beta * (L @ beta)))


def auto_tune_lambda(Y_sketch, X_sketch: np.ndarray, A, alpha: float = 0.005) -> float:
    scale = float(np.mean(np.einsum("kd,kd->k", X_sketch, X_sketch)))
    mean_deg = float(np.mean(np.asarray(A.sum(axis=1)).ravel())) if A.shape[0] else 0.0
    return float(alpha * scale / max(mean_deg, 1.0))
