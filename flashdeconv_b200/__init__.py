"""flashdeconv_b200 -- B200-native hot path of FlashDeconv behind the reference's API.

    from flashdeconv_b200 import FlashDeconv
    proportions = FlashDeconv(sketch_dim=512).fit_transform(Y, X, coords)

Importing the package does not touch CUDA; the native library (libfdb200.so,
built by ``python -m flashdeconv_b200.build``) is loaded on first use and there
is no CPU fallback.
"""
__version__ = "0.1.0"

from .estimator import FlashDeconv
from . import tl
from . import io
from . import core, utils          # the reference's sub-package import paths

__all__ = ["FlashDeconv", "tl", "__version__"]
