"""`fd.tl.deconvolve` over the B200 path.

Keeps the reference signature (tl/_deconvolve.py:6-26) and the fields it writes
(.obsm[key], .obs[key+"_dominant"], .uns[key+"_params"]).  The AnnData plumbing
of the reference's io/loader.py is file/frame handling, not part of the
accelerated path; the few lines needed here are restated inline and only need
an AnnData-like object (``.X``/``.layers``, ``.obs``, ``.var_names``, ``.obsm``, ``.uns``).
"""
from __future__ import annotations

from typing import Any, Optional, Union

import numpy as np
from scipy import sparse


def _matrix(adata, layer):
    return adata.X if layer is None else adata.layers[layer]


def _signatures(adata_ref, cell_type_key, layer):
    """Mean expression per cell type (io/loader.py:119-136).  A sparse cells x genes matrix is reduced on the device
    (one pass, float64 accumulation); dense references take numpy's grouped mean."""
    M = _matrix(adata_ref, layer)
    labels = np.asarray(adata_ref.obs[cell_type_key])
    names, codes = np.unique(labels, return_inverse=True)
    if sparse.issparse(M) and (M.dtype == np.float32 or np.issubdtype(M.dtype, np.integer)):    # exact in float32
        from .. import pipeline
        return pipeline.group_means(pipeline.csr_to_device(M), codes, len(names)), names
    M = np.asarray(M)
    X = np.zeros((len(names), M.shape[1]), dtype=np.float64)
    for r in range(len(names)):
        X[r] = M[codes == r].mean(axis=0)
    return X, names


def deconvolve(adata_st: Any, adata_ref: Any, cell_type_key: str = "cell_type", *, sketch_dim: int = 512,
               lambda_spatial: Union[float, str] = "auto", rho_sparsity: float = 0.01, n_hvg: int = 2000,
               n_markers_per_type: int = 50, spatial_method: str = "knn", k_neighbors: int = 6,
               radius: Optional[float] = None, preprocess: str = "log_cpm", layer_st: Optional[str] = None,
               layer_ref: Optional[str] = None, spatial_key: str = "spatial", key_added: str = "flashdeconv",
               random_state: int = 0, copy: bool = False) -> Optional[Any]:
    from ..estimator import FlashDeconv

    adata = adata_st.copy() if copy else adata_st
    if spatial_key not in adata.obsm:
        raise ValueError(f"Spatial coordinates not found in adata.obsm['{spatial_key}']")
    if cell_type_key not in adata_ref.obs:
        raise ValueError(f"Cell type key '{cell_type_key}' not found in adata_ref.obs")
    Y = _matrix(adata, layer_st)
    coords = np.asarray(adata.obsm[spatial_key], dtype=np.float64)
    X, names = _signatures(adata_ref, cell_type_key, layer_ref)
    st_genes, ref_genes = np.asarray(adata.var_names), np.asarray(adata_ref.var_names)
    common, i_st, i_ref = np.intersect1d(st_genes, ref_genes, return_indices=True)
    if len(common) == 0:
        raise ValueError("No common genes found between spatial and reference data")
    Y = Y[:, i_st]
    if sparse.issparse(Y):
        Y = Y.tocsr()
    X = X[:, i_ref]

    model = FlashDeconv(sketch_dim=sketch_dim, lambda_spatial=lambda_spatial, rho_sparsity=rho_sparsity,
                        n_hvg=n_hvg, n_markers_per_type=n_markers_per_type, spatial_method=spatial_method,
                        k_neighbors=k_neighbors, radius=radius, preprocess=preprocess,
                        random_state=random_state, verbose=False)
    proportions = model.fit_transform(Y, X, coords, cell_type_names=names)

    try:
        import pandas as pd
        adata.obsm[key_added] = pd.DataFrame(proportions, index=adata.obs_names, columns=list(names))
        adata.obs[f"{key_added}_dominant"] = pd.Categorical(np.asarray(names)[model.get_dominant_cell_type()],
                                                            categories=list(names))
    except ImportError:                                   # pragma: no cover
        adata.obsm[key_added] = proportions
        adata.obs[f"{key_added}_dominant"] = np.asarray(names)[model.get_dominant_cell_type()]
    adata.uns[f"{key_added}_params"] = {
        "sketch_dim": sketch_dim, "lambda_spatial": float(model.lambda_used_), "rho_sparsity": rho_sparsity,
        "n_hvg": n_hvg, "n_markers_per_type": n_markers_per_type, "spatial_method": spatial_method,
        "k_neighbors": k_neighbors, "radius": radius, "preprocess": preprocess,
        "n_genes_used": len(model.gene_idx_), "n_cell_types": len(names), "cell_type_names": list(names),
        "random_state": random_state, "converged": model.info_.get("converged", False),
        "n_iterations": model.info_.get("n_iterations", 0),
    }
    return adata if copy else None
