"""``flashdeconv.core.sketching`` import path: the mirror lives in ``flashdeconv_b200.sketching``."""
from ..sketching import build_countsketch_matrix, project_to_sketch, sketch_data          # noqa: F401
