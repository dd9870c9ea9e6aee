"""`FlashDeconv`: the reference's estimator surface (core/deconv.py:20-512) over the B200 path.

Same constructor, validation messages, methods and fitted attributes as the
reference class; steps 2-6 of ``fit`` run on the GPU through libfdb200 (there is
no CPU fallback).  All three ``preprocess`` values run on the device: "log_cpm"
transforms values inside the fused sketch kernel; "raw" and "pearson" are per-gene
scalings folded into the device-side gene weights (core/deconv.py:199-229).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Union

import numpy as np
from scipy import sparse

from . import genes


class FlashDeconv:
    def __init__(self, sketch_dim: int = 512, lambda_spatial: Union[float, str] = "auto",
                 rho_sparsity: float = 0.01, n_hvg: int = 2000, n_markers_per_type: int = 50,
                 spatial_method: str = "knn", k_neighbors: int = 6, radius: Optional[float] = None,
                 max_iter: int = 100, tol: float = 1e-4, preprocess: str = "log_cpm",
                 random_state: Optional[int] = 0, verbose: bool = False):
        checks = [
            (sketch_dim <= 0, f"sketch_dim must be positive, got {sketch_dim}"),
            (k_neighbors < 0, f"k_neighbors must be non-negative, got {k_neighbors}"),
            (max_iter < 0, f"max_iter must be non-negative, got {max_iter}"),
            (tol <= 0, f"tol must be positive, got {tol}"),
            (isinstance(lambda_spatial, (int, float)) and lambda_spatial < 0,
             f"lambda_spatial must be non-negative, got {lambda_spatial}"),
            (rho_sparsity < 0, f"rho_sparsity must be non-negative, got {rho_sparsity}"),
            (n_hvg < 0, f"n_hvg must be non-negative, got {n_hvg}"),
            (n_markers_per_type < 0, f"n_markers_per_type must be non-negative, got {n_markers_per_type}"),
            (spatial_method == "radius" and radius is None,
             "radius must be specified when spatial_method='radius'"),
            (radius is not None and radius <= 0, f"radius must be positive, got {radius}"),
        ]
        for bad, msg in checks:
            if bad:
                raise ValueError(msg)
        self.sketch_dim = sketch_dim
        self.lambda_spatial = lambda_spatial
        self.rho_sparsity = rho_sparsity
        self.n_hvg = n_hvg
        self.n_markers_per_type = n_markers_per_type
        self.spatial_method = spatial_method
        self.k_neighbors = k_neighbors
        self.radius = radius
        self.max_iter = max_iter
        self.tol = tol
        self.preprocess = preprocess
        self.random_state = random_state
        self.verbose = verbose
        self.beta_ = None
        self.proportions_ = None
        self.gene_idx_ = None
        self.info_ = None
        self._fitted = False
        self._graph = None
        self._adjacency = None

    # adjacency_ is materialised (device -> scipy CSR, input order) on first access
    @property
    def adjacency_(self):
        if self._adjacency is None and self._graph is not None:
            self._adjacency = self._graph.to_scipy()
        return self._adjacency

    def fit(self, Y, X: np.ndarray, coords: np.ndarray, cell_type_names: Optional[np.ndarray] = None):
        if Y.shape[1] != X.shape[1]:
            raise ValueError(f"Gene dimension mismatch: Y has {Y.shape[1]} genes but X has {X.shape[1]} genes. "
                             "They must share the same gene space (align before calling fit).")
        if coords.shape[0] != Y.shape[0]:
            raise ValueError(f"Spot count mismatch: Y has {Y.shape[0]} spots but coords has {coords.shape[0]} rows. "
                             "Each spot needs exactly one coordinate.")
        if X.shape[0] == 0:
            raise ValueError("Reference matrix X must contain at least one cell type (X.shape[0] > 0). "
                             "Check your reference filtering and cell_type_key mapping.")
        if cell_type_names is not None and len(cell_type_names) != X.shape[0]:
            raise ValueError(f"cell_type_names length ({len(cell_type_names)}) does not match number of "
                             f"cell types in X ({X.shape[0]}).")
        if self.preprocess not in ("log_cpm", "pearson", "raw"):
            raise ValueError(f"Unknown preprocess method: {self.preprocess}. "
                             "Choose from 'log_cpm', 'pearson', or 'raw'.")
        from . import pipeline        # imports the native library: fails loudly when it is missing
        if X.shape[0] > pipeline.MAX_TYPES_WIDE:
            raise ValueError(f"flashdeconv_b200 supports at most {pipeline.MAX_TYPES_WIDE} cell types "
                             f"(register-resident kernels up to {pipeline.MAX_TYPES}, warp-per-spot kernels beyond); "
                             f"got {X.shape[0]}.")

        say = print if self.verbose else (lambda *a, **k: None)
        say("FlashDeconv: Starting deconvolution...")
        say(f"  Spatial data: {Y.shape[0]} spots x {Y.shape[1]} genes")
        say(f"  Reference: {X.shape[0]} cell types x {X.shape[1]} genes")
        self.n_spots_, self.n_genes_ = Y.shape
        self.n_cell_types_ = X.shape[0]
        self.cell_type_names_ = cell_type_names

        pipeline._native.require_cuda()
        say("Step 1: Selecting informative genes...")
        # counts go to the device once; the O(nnz) moment pass of the HVG selection runs there too
        # (SURVEY 8 f1), the G-sized ranking and the K x G_sel SVD stay on the host
        csr = pipeline.csr_to_device(Y)
        gene_idx, leverage = genes.select_informative_genes_device(csr, np.asarray(X), n_hvg=self.n_hvg,
                                                                   n_markers_per_type=self.n_markers_per_type)
        self.gene_idx_ = gene_idx
        say(f"  Selected {len(gene_idx)} genes (HVG + markers)")
        say(f"Step 2-3: {self.preprocess} preprocessing + sketching to {self.sketch_dim} dimensions (fused, on device)...")
        say("Step 4-6: spatial graph, lambda, block coordinate descent...")
        res = pipeline.deconvolve_path(csr, X, coords, gene_idx, leverage, sketch_dim=self.sketch_dim,
                                       lambda_spatial=self.lambda_spatial, rho_sparsity=self.rho_sparsity,
                                       spatial_method=self.spatial_method, k_neighbors=self.k_neighbors,
                                       radius=self.radius, max_iter=self.max_iter, tol=self.tol,
                                       random_state=self.random_state, verbose=self.verbose,
                                       preprocess=self.preprocess, pinned_out=True)   # outputs land in page-locked arrays
        self._graph, self._adjacency = res.graph, None
        self.lambda_used_ = res.lambda_used
        self.beta_, self.proportions_, self.info_ = res.beta, res.proportions, res.info
        self._dominant = res.dominant
        self._fitted = True
        n = max(Y.shape[0], 1)
        say(f"  Average neighbors per spot: {res.graph.nnz / n:.1f}")
        say(f"  lambda = {res.lambda_used:.4f}")
        say(f"  Converged: {res.info['converged']}")
        say(f"  Iterations: {res.info['n_iterations']}")
        say("FlashDeconv: Done!")
        return self

    def fit_transform(self, Y, X: np.ndarray, coords: np.ndarray, **kwargs) -> np.ndarray:
        self.fit(Y, X, coords, **kwargs)
        return self.proportions_

    def _need_fit(self):
        if not self._fitted:
            raise RuntimeError("Model has not been fitted. Call fit() first.")

    def get_cell_type_proportions(self) -> np.ndarray:
        self._need_fit()
        return self.proportions_

    def get_abundances(self) -> np.ndarray:
        self._need_fit()
        return self.beta_

    def get_dominant_cell_type(self) -> np.ndarray:
        """argmax over the cell types, computed on the device with the solve (core/deconv.py:467-478)."""
        self._need_fit()
        if getattr(self, "_dominant", None) is not None:
            return self._dominant.astype(np.intp)
        return np.argmax(self.proportions_, axis=1)

    def summary(self) -> Dict[str, Any]:
        if not self._fitted:
            return {"fitted": False}
        return {"fitted": True, "n_spots": self.n_spots_, "n_cell_types": self.n_cell_types_,
                "n_genes_used": len(self.gene_idx_), "sketch_dim": self.sketch_dim,
                "lambda_spatial": self.lambda_used_, "rho_sparsity": self.rho_sparsity,
                "preprocess_method": self.preprocess, "converged": self.info_["converged"],
                "n_iterations": self.info_["n_iterations"], "final_objective": self.info_["final_objective"]}

    def __repr__(self) -> str:
        return (f"FlashDeconv(sketch_dim={self.sketch_dim}, lambda_spatial={self.lambda_spatial}, "
                f"status={'fitted' if self._fitted else 'not fitted'})")
