"""microaligner_b200 -- B200-native non-linear registration hot path of microaligner.

    from microaligner_b200 import OptFlowRegistrator, Warper, pad_to_shape, transform_img_with_tmat

mirrors `from microaligner import OptFlowRegistrator, Warper, pad_to_shape, transform_img_with_tmat`
(reference microaligner/__init__.py:19-20; FeatureRegistrator, the sparse-feature matcher, is out of scope).
Importing this package loads libmicroaligner_b200.so and fails loudly if it is missing."""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA extension is not built)
from .optflow_reg import OptFlowRegistrator, Warper
from .shared_modules.utils import pad_to_shape, transform_img_with_tmat

__all__ = ["OptFlowRegistrator", "Warper", "pad_to_shape", "transform_img_with_tmat"]
