"""``flashdeconv.utils.genes`` import path: the mirror lives in ``flashdeconv_b200.genes``."""
from ..genes import compute_leverage_scores, select_hvg, select_informative_genes, select_markers          # noqa: F401
