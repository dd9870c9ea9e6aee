"""Host-side data hand-off around the device path: the five helpers of the reference's ``flashdeconv/io/loader.py`` with the
same names, arguments, return values and error messages, working on AnnData-like objects (``.X`` / ``.layers``, ``.obs``,
``.obsm``, ``.var_names``, ``.obs_names``, ``.n_obs``).  The only part with real work in it -- the per-cell-type aggregation
of a single-cell reference (io/loader.py:119-136) -- runs on the device for sparse count matrices
(``pipeline.group_means``, float64 accumulation, one pass over the CSR); everything else is index bookkeeping.
"""
from __future__ import annotations

from typing import Any, Optional, Tuple, Union

import numpy as np
from scipy import sparse

ArrayLike = Union[np.ndarray, sparse.spmatrix]


def _expression(adata, layer):
    return adata.X if layer is None else adata.layers[layer]


def load_spatial_data(adata: Any, layer: Optional[str] = None, coord_key: str = "spatial") -> Tuple[ArrayLike, np.ndarray, np.ndarray]:
    """(Y, coords, gene_names) of a spatial AnnData (io/loader.py:15-70).  Coordinates are looked up in
    ``.obsm[coord_key]``, ``.obsm["X_spatial"]``, ``.obs["x"], ["y"]``, ``.obs["array_row"], ["array_col"]``, in that order."""
    Y = _expression(adata, layer)
    if coord_key in adata.obsm:
        coords = np.array(adata.obsm[coord_key])
    elif "X_spatial" in adata.obsm:
        coords = np.array(adata.obsm["X_spatial"])
    elif "x" in adata.obs and "y" in adata.obs:
        coords = np.column_stack([adata.obs["x"], adata.obs["y"]])
    elif "array_row" in adata.obs and "array_col" in adata.obs:
        coords = np.column_stack([adata.obs["array_row"], adata.obs["array_col"]])
    else:
        raise ValueError(f"Could not find spatial coordinates. Expected key '{coord_key}' in adata.obsm or 'x'/'y' in adata.obs")
    return Y, coords, np.array(adata.var_names)


def load_reference(adata_ref: Any, cell_type_key: str = "cell_type", layer: Optional[str] = None,
                   method: str = "mean") -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(X, cell_type_names, gene_names): one row per cell type, the mean (or sum) of its cells (io/loader.py:73-140)."""
    expr = _expression(adata_ref, layer)
    if cell_type_key not in adata_ref.obs:
        raise ValueError(f"Cell type key '{cell_type_key}' not found in adata_ref.obs")
    if method not in ("mean", "sum"):
        raise ValueError(f"Unknown aggregation method: {method}")
    labels = np.array(adata_ref.obs[cell_type_key])
    names, codes = np.unique(labels, return_inverse=True)
    counts = np.bincount(codes, minlength=len(names)).astype(np.float64)
    if sparse.issparse(expr) and (expr.dtype == np.float32 or np.issubdtype(expr.dtype, np.integer)):
        # sparse counts (exact in float32): grouped sums on the device, float64 accumulation
        from .. import pipeline
        X = pipeline.group_means(pipeline.csr_to_device(expr), codes, len(names))
        if method == "sum":
            X = X * counts[:, None]
    else:                                                      # dense or float64 input: numpy, as the reference
        X = np.zeros((len(names), expr.shape[1]), dtype=np.float64)
        for i in range(len(names)):
            rows = expr[codes == i]
            total = np.asarray(rows.sum(axis=0)).ravel() if sparse.issparse(expr) else np.sum(rows, axis=0)
            X[i] = total / counts[i] if method == "mean" else total
    return X, names, np.array(adata_ref.var_names)


def align_genes(Y: ArrayLike, X: np.ndarray, genes_spatial: np.ndarray, genes_ref: np.ndarray) -> Tuple[ArrayLike, np.ndarray, np.ndarray]:
    """Restricts both matrices to the genes they share, in sorted name order; a repeated name refers to its first
    occurrence (io/loader.py:143-194)."""
    genes_spatial, genes_ref = np.asarray(genes_spatial), np.asarray(genes_ref)
    common = np.intersect1d(genes_spatial, genes_ref)
    if len(common) == 0:
        raise ValueError("No common genes found between spatial data and reference")

    def first_positions(names):
        uniq, first = np.unique(names, return_index=True)      # index of the first occurrence of every name
        return first[np.searchsorted(uniq, common)]

    return Y[:, first_positions(genes_spatial)], X[:, first_positions(genes_ref)], common


def result_to_anndata(beta: np.ndarray, adata: Any, cell_type_names: Optional[np.ndarray] = None,
                      key_added: str = "flashdeconv") -> Any:
    """Stores an (n_spots, n_cell_types) result as a DataFrame in ``adata.obsm[key_added]`` and the dominant type as a
    categorical in ``adata.obs[key_added + "_dominant"]`` (io/loader.py:197-258)."""
    import pandas as pd
    beta = np.asarray(beta)
    if beta.ndim != 2:
        raise ValueError(f"beta must be 2D, got shape {beta.shape}")
    if beta.shape[0] != adata.n_obs:
        raise ValueError(f"beta rows must match adata.n_obs, got beta.shape[0]={beta.shape[0]} and adata.n_obs={adata.n_obs}")
    columns = np.asarray(cell_type_names) if cell_type_names is not None else \
        np.array([f"CellType_{i}" for i in range(beta.shape[1])])
    if len(columns) != beta.shape[1]:
        raise ValueError(f"Length of cell_type_names ({len(columns)}) must match beta.shape[1] ({beta.shape[1]})")
    adata.obsm[key_added] = pd.DataFrame(beta, index=adata.obs_names, columns=columns)
    adata.obs[f"{key_added}_dominant"] = pd.Categorical(columns[np.argmax(beta, axis=1)], categories=columns)
    return adata


def prepare_data(adata_st: Any, adata_ref: Any, cell_type_key: str = "cell_type", spatial_coord_key: str = "spatial",
                 layer_st: Optional[str] = None, layer_ref: Optional[str] = None):
    """(Y, X, coords, cell_type_names, gene_names) ready for ``FlashDeconv.fit`` (io/loader.py:261-318)."""
    Y, coords, genes_st = load_spatial_data(adata_st, layer=layer_st, coord_key=spatial_coord_key)
    X, names, genes_ref = load_reference(adata_ref, cell_type_key=cell_type_key, layer=layer_ref)
    Y, X, genes = align_genes(Y, X, genes_st, genes_ref)
    return Y, X, coords, names, genes
