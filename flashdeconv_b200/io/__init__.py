"""AnnData hand-off helpers with the reference's names (flashdeconv/io/__init__.py:3-17)."""
from .loader import align_genes, load_reference, load_spatial_data, prepare_data, result_to_anndata

__all__ = ["load_spatial_data", "load_reference", "align_genes", "result_to_anndata", "prepare_data"]
