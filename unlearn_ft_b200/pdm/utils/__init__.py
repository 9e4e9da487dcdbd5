from .metric_utils import compute_snr  # noqa: F401
from .estimation_utils import hard_concrete, keep_indices  # noqa: F401
