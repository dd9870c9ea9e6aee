"""``flashdeconv.core.solver`` import path: the mirror lives in ``flashdeconv_b200.solver``."""
from ..solver import (bcd_solve, compute_objective, normalize_proportions, precompute_gram_matrix,          # noqa: F401
                      soft_threshold)
