"""``flashdeconv.core.deconv`` import path: the estimator lives in ``flashdeconv_b200.estimator``."""
from ..estimator import FlashDeconv          # noqa: F401
