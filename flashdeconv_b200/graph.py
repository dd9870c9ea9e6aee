"""Function-level mirror of the reference's ``flashdeconv/utils/graph.py`` over libfdb200.

``build_knn_graph`` (:25-83), ``build_radius_graph`` (:86-133), ``build_grid_graph``
(:136-172) and ``coords_to_adjacency`` (:175-212) return scipy CSR float64 binary
adjacency in input order with ascending columns, built by the grid-hash kernels.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
from scipy import sparse


def _validate_coords(coords: np.ndarray) -> None:
    if coords.ndim != 2 or coords.shape[1] == 0:
        raise ValueError("coords must be 2D with at least 1 coordinate dimension, "
                         f"got shape {coords.shape}")


def _device_graph(coords, method, k=6, radius=None):
    from . import pipeline
    torch = pipeline._native.require_cuda()
    c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).cuda()
    return pipeline.build_graph(c, method, k, radius)


def _with_self(A, n, include_self):
    if include_self and n > 0:
        A = (A + sparse.eye(n, dtype=np.float64, format="csr")).tocsr()
        A.data[:] = 1.0
    return A


def build_knn_graph(coords: np.ndarray, k: int = 6, include_self: bool = False) -> sparse.csr_matrix:
    _validate_coords(coords)
    n = coords.shape[0]
    if min(k, n - 1) <= 0:
        return _with_self(sparse.csr_matrix((n, n), dtype=np.float64), n, include_self)
    return _with_self(_device_graph(coords, "knn", k=k).to_scipy(), n, include_self)


def build_radius_graph(coords: np.ndarray, radius: float, include_self: bool = False) -> sparse.csr_matrix:
    _validate_coords(coords)
    n = coords.shape[0]
    if n == 0:
        return sparse.csr_matrix((0, 0), dtype=np.float64)
    return _with_self(_device_graph(coords, "radius", radius=radius).to_scipy(), n, include_self)


def build_grid_graph(coords: np.ndarray, grid_spacing: Optional[float] = None) -> sparse.csr_matrix:
    _validate_coords(coords)
    n = coords.shape[0]
    if n <= 1:
        return sparse.csr_matrix((n, n), dtype=np.float64)
    if grid_spacing is not None:
        return build_radius_graph(coords, grid_spacing * 1.5)
    return _device_graph(coords, "grid").to_scipy()


def coords_to_adjacency(coords: np.ndarray, method: str = "knn", k: int = 6,
                        radius: Optional[float] = None) -> sparse.csr_matrix:
    if method == "knn":
        return build_knn_graph(coords, k=k)
    if method == "radius":
        if radius is None:
            raise ValueError("radius must be specified for radius method")
        return build_radius_graph(coords, radius=radius)
    if method == "grid":
        return build_grid_graph(coords)
    raise ValueError(f"Unknown method: {method}")
