"""Type aliases kept under the reference's names (shared_modules/dtype_aliases.py:23-42)."""
from typing import Tuple, Union

import numpy as np
import torch

Image = Union[np.ndarray, torch.Tensor]   # 2-D grey image, host or device
Flow = Union[np.ndarray, torch.Tensor]    # (H, W, 2) float32 optical-flow map
Shape2D = Tuple[int, int]
