"""``flashdeconv.core.spatial`` import path: the mirror lives in ``flashdeconv_b200.spatial``."""
from ..spatial import (auto_tune_lambda, compute_degree_matrix, compute_laplacian, compute_laplacian_quadratic,   # noqa: F401
                       get_neighbor_counts, get_neighbor_indices)
