"""gpar_b200 -- B200-native (sm_100a) engine for the per-layer GP hot path of GPAR.

Public API mirrors wesselb/gpar: :class:`GPARRegressor` (gpar/regression.py) and
:class:`GPAR` (gpar/model.py).  Importing the package needs neither a GPU nor the
compiled library; using it does (there is no CPU fallback)."""
from .model import GPAR  # noqa: F401
from .regression import GPARRegressor, log_transform, squishing_transform  # noqa: F401

__all__ = ["GPAR", "GPARRegressor", "log_transform", "squishing_transform"]
