"""The reference's ``flashdeconv.utils`` import path (utils/__init__.py:3-34) over the mirrors.  The evaluation metrics of
``utils/metrics.py`` are not part of the accelerated path and are not provided."""
from ..genes import compute_leverage_scores, select_hvg, select_markers
from ..graph import build_knn_graph, build_radius_graph, coords_to_adjacency
from .random import check_random_state
from . import genes, graph, random

__all__ = ["select_hvg", "select_markers", "compute_leverage_scores", "build_knn_graph", "build_radius_graph",
           "coords_to_adjacency", "check_random_state"]
