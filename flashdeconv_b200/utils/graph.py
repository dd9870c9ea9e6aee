"""``flashdeconv.utils.graph`` import path: the mirror lives in ``flashdeconv_b200.graph``."""
from ..graph import build_grid_graph, build_knn_graph, build_radius_graph, coords_to_adjacency          # noqa: F401
