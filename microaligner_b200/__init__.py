"""microaligner_b200 -- B200-native non-linear registration hot path of microaligner.

    from microaligner_b200 import OptFlowRegistrator, Warper

mirrors `from microaligner import OptFlowRegistrator, Warper` (reference microaligner/__init__.py:19).
Importing this package loads libmicroaligner_b200.so and fails loudly if it is missing."""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA extension is not built)
from .optflow_reg import OptFlowRegistrator, Warper

__all__ = ["OptFlowRegistrator", "Warper"]
