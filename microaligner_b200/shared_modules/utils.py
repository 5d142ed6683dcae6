"""pad_to_shape / transform_img_with_tmat -- the two helpers the reference exports next to its registrators
(microaligner/__init__.py:20; shared_modules/utils.py:40-66, 98-114) and applies to every page of a cycle in the
linear-registration branch of the pipeline (transform_and_save_zplanes, __main__.py:84-112).

Same signatures and results, computed on the device: padding is index math inside the resampling kernel
(ma_warp_affine), and the resampling reproduces skimage.transform.warp(order=1, mode='constant', cval=0,
preserve_range=True) in float64.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.

scikit-image is not part of this environment: parity of the resampling is pinned against oracle/affine_np.py, a
restatement of skimage's published algorithm, NOT against skimage itself."""
from typing import Tuple, Union

import numpy as np
import torch

from .. import ops

Image = Union[np.ndarray, torch.Tensor]
Padding = Tuple[int, int, int, int]

_IDENTITY = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
_EYE3 = np.eye(3)


def _calculate_padding_size(bigger_shape: int, smaller_shape: int) -> Tuple[int, int]:
    """Split the size difference in two; the odd pixel goes to the second (right / bottom) side (utils.py:40-50)."""
    diff = bigger_shape - smaller_shape
    if diff == 1:
        return 0, 1
    if diff % 2 != 0:
        return int(diff // 2), int(diff // 2 + 1)
    return int(diff / 2), int(diff / 2)


def _padding(shape, target_shape) -> Padding:
    left, right = _calculate_padding_size(target_shape[1], shape[1])
    top, bottom = _calculate_padding_size(target_shape[0], shape[0])
    if min(left, right, top, bottom) < 0:
        raise ValueError(f"cannot pad an image of shape {tuple(shape)} to the smaller shape {tuple(target_shape)}")
    return left, right, top, bottom


def pad_to_shape(img: Image, target_shape: Tuple[int, int]) -> Tuple[Image, Padding]:
    """Zero-pad `img` to target_shape, centred; returns (padded image, (left, right, top, bottom))."""
    target_shape = (int(target_shape[0]), int(target_shape[1]))
    if tuple(img.shape) == target_shape:
        return img, (0, 0, 0, 0)
    pad = _padding(img.shape, target_shape)
    dev = ops.to_device(img)
    out = ops.warp_affine(dev, _EYE3, target_shape, pad_top=pad[2], pad_left=pad[0])   # identity map = exact copy
    return (out if isinstance(img, torch.Tensor) else ops.to_host(out)), pad


def transform_img_with_tmat(img: Image, target_shape: Tuple[int, int], transform_matrix: np.ndarray) -> Image:
    """Pad to target_shape, then resample by the 2x3 matrix `transform_matrix` (moving -> reference coordinates).
    The identity matrix only pads, as in the reference."""
    target_shape = (int(target_shape[0]), int(target_shape[1]))
    tmat = np.asarray(transform_matrix, dtype=np.float64)
    if tmat.shape != (2, 3):
        raise ValueError("transform_matrix must be a 2x3 affine matrix")
    if np.array_equal(tmat, _IDENTITY):
        return pad_to_shape(img, target_shape)[0]
    # partial inverse, to survive singular matrices (utils.py:106-108)
    inv_matrix = np.linalg.pinv(np.append(tmat, [[0, 0, 1]], axis=0))
    pad = (0, 0, 0, 0) if tuple(img.shape) == target_shape else _padding(img.shape, target_shape)
    out = ops.warp_affine(ops.to_device(img), inv_matrix, target_shape, pad_top=pad[2], pad_left=pad[0])
    return out if isinstance(img, torch.Tensor) else ops.to_host(out)
