"""The reference's ``flashdeconv.core`` import path (core/__init__.py:3-22) over the device-backed mirrors."""
from ..estimator import FlashDeconv
from ..sketching import build_countsketch_matrix, project_to_sketch
from ..solver import bcd_solve
from ..spatial import compute_laplacian, get_neighbor_indices
from . import deconv, sketching, solver, spatial

__all__ = ["FlashDeconv", "build_countsketch_matrix", "project_to_sketch", "compute_laplacian", "get_neighbor_indices",
           "bcd_solve"]
