"""scanpy-style entry point (reference: flashdeconv/tl/_deconvolve.py:6-174)."""
from ._deconvolve import deconvolve

__all__ = ["deconvolve"]
