"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's affine page transform.

    transform_img_with_tmat(img, target_shape, tmat)      reference shared_modules/utils.py:98-114
    pad_to_shape / _calculate_padding_size                reference shared_modules/utils.py:40-66

The arithmetic lives in scikit-image (`skimage.transform.warp` with an `AffineTransform`, pinned
scikit-image==0.19.2 in the reference's environment.yaml), which is NOT installed in this image and is
not vendored under /root/reference.

    *** PARITY UNPINNED against scikit-image itself. ***

What follows restates the published algorithm of skimage/transform/_warps.py::warp (order=1,
mode='constant', cval=0, clip=True, preserve_range=True) -> _warps_cy.pyx::_warp_fast ->
interpolation.pxd::bilinear_interpolation, all in float64 for integer input:

  * the 3x3 matrix handed to warp is  pinv([[tmat],[0,0,1]])  (utils.py:106-109), used as the
    output->input map; _warp_fast picks the coordinate transform from the matrix' last row:
       M[2]==(0,0,1) and M[0,1]==M[1,0]==0 :  c = M00*x + M02          r = M11*y + M12
       M[2]==(0,0,1)                        :  c = M00*x + M01*y + M02  r = M10*x + M11*y + M12
       otherwise (projective)               :  z = M20*x + M21*y + M22, c = (...)/z, r = (...)/z
    with x = output column, y = output row, products and sums rounded separately, left to right;
  * bilinear_interpolation: minr=floor(r), maxr=ceil(r), dr=r-minr (same for c); pixels outside the
    image are cval=0;  top=(1-dc)*tl+dc*tr, bottom=(1-dc)*bl+dc*br, out=(1-dr)*top+dr*bottom;
  * the clip to [min(img.min(),0), max(img.max(),0)] cannot change the integer part of a convex
    combination of in-range values, and `.astype(original_dtype)` truncates towards zero.

Cross-checks available here (tests/test_oracle_affine.py): integer translations are exact shifts,
the identity matrix returns the padded input, and scipy.ndimage.affine_transform(order=1) -- a
different implementation of the same interpolant -- agrees to +-1 grey level."""
from typing import Tuple

import numpy as np

IDENTITY = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])


def calculate_padding_size(bigger: int, smaller: int) -> Tuple[int, int]:
    """utils.py:40-50."""
    diff = bigger - smaller
    if diff == 1:
        return 0, 1
    if diff % 2 != 0:
        return int(diff // 2), int(diff // 2 + 1)
    return int(diff / 2), int(diff / 2)


def pad_to_shape(img: np.ndarray, target_shape: Tuple[int, int]):
    """utils.py:53-66: zero padding, centred (the extra pixel goes right / bottom)."""
    if tuple(img.shape) == tuple(target_shape):
        return img, (0, 0, 0, 0)
    left, right = calculate_padding_size(target_shape[1], img.shape[1])
    top, bottom = calculate_padding_size(target_shape[0], img.shape[0])
    if min(left, right, top, bottom) < 0:
        raise ValueError("target shape is smaller than the image")   # cv.copyMakeBorder raises as well
    out = np.zeros(target_shape, img.dtype)
    out[top:top + img.shape[0], left:left + img.shape[1]] = img
    return out, (left, right, top, bottom)


def inverse_matrix(tmat: np.ndarray) -> np.ndarray:
    """utils.py:106-108: pseudo-inverse of the 3x3 homogeneous matrix (handles singular input)."""
    m3 = np.append(np.asarray(tmat, dtype=np.float64), [[0, 0, 1]], axis=0)
    return np.linalg.pinv(m3)


def transform_kind(M: np.ndarray) -> int:
    """_warp_fast's dispatch: 0 metric, 1 affine, 2 projective."""
    if M[2, 0] == 0 and M[2, 1] == 0 and M[2, 2] == 1:
        return 0 if (M[0, 1] == 0 and M[1, 0] == 0) else 1
    return 2


def warp_fast_bilinear(img: np.ndarray, M: np.ndarray) -> np.ndarray:
    """_warp_fast(order=1, mode='constant', cval=0) on a float64 copy of img; float64 result."""
    rows, cols = img.shape
    f = img.astype(np.float64)
    x = np.arange(cols, dtype=np.float64)[None, :]
    y = np.arange(rows, dtype=np.float64)[:, None]
    kind = transform_kind(M)
    with np.errstate(all="ignore"):
        if kind == 0:
            c = np.broadcast_to(M[0, 0] * x + M[0, 2], (rows, cols))
            r = np.broadcast_to(M[1, 1] * y + M[1, 2], (rows, cols))
        elif kind == 1:
            c = (M[0, 0] * x + M[0, 1] * y) + M[0, 2]
            r = (M[1, 0] * x + M[1, 1] * y) + M[1, 2]
        else:
            z = (M[2, 0] * x + M[2, 1] * y) + M[2, 2]
            c = ((M[0, 0] * x + M[0, 1] * y) + M[0, 2]) / z
            r = ((M[1, 0] * x + M[1, 1] * y) + M[1, 2]) / z
        ok = np.isfinite(r) & np.isfinite(c) & (np.abs(r) < 2.0 ** 31) & (np.abs(c) < 2.0 ** 31)
        r = np.where(ok, r, -10.0)      # anything that cannot touch the image reads four zeros -> 0
        c = np.where(ok, c, -10.0)
        minr, minc = np.floor(r), np.floor(c)
        maxr, maxc = np.ceil(r), np.ceil(c)
        dr, dc = r - minr, c - minc

        def px(rr, cc):
            inside = (rr >= 0) & (rr < rows) & (cc >= 0) & (cc < cols)
            ri = np.clip(rr, 0, rows - 1).astype(np.int64)
            ci = np.clip(cc, 0, cols - 1).astype(np.int64)
            return np.where(inside, f[ri, ci], 0.0)

        top = (1 - dc) * px(minr, minc) + dc * px(minr, maxc)
        bottom = (1 - dc) * px(maxr, minc) + dc * px(maxr, maxc)
        return (1 - dr) * top + dr * bottom


def transform_img_with_tmat(img: np.ndarray, target_shape: Tuple[int, int], transform_matrix: np.ndarray) -> np.ndarray:
    """utils.py:98-114."""
    dtype = img.dtype
    img, _ = pad_to_shape(img, tuple(target_shape))
    if np.array_equal(transform_matrix, IDENTITY):
        return img
    out = warp_fast_bilinear(img, inverse_matrix(transform_matrix))
    lo, hi = min(float(img.min()), 0.0), max(float(img.max()), 0.0)
    return np.clip(out, lo, hi).astype(dtype)
