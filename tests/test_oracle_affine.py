"""CPU checks of oracle/affine_np.py (the restatement of transform_img_with_tmat, utils.py:98-114).
scikit-image is not installed here, so the oracle is cross-checked against properties of the interpolant
and against scipy.ndimage.affine_transform -- an independent implementation of order-1 interpolation."""
import numpy as np
import pytest

from oracle import affine_np as A


def image(shape=(97, 131), dtype=np.uint16, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, np.iinfo(dtype).max, shape, endpoint=True).astype(dtype)


@pytest.mark.parametrize("big,small,expect", [(10, 10, (0, 0)), (11, 10, (0, 1)), (13, 10, (1, 2)), (14, 10, (2, 2))])
def test_padding_sizes(big, small, expect):
    assert A.calculate_padding_size(big, small) == expect


def test_pad_to_shape_matches_copy_make_border():
    cv2 = pytest.importorskip("cv2")
    img = image((30, 41))
    out, (l, r, t, b) = A.pad_to_shape(img, (37, 44))
    assert (l, r, t, b) == (1, 2, 3, 4)
    assert np.array_equal(out, cv2.copyMakeBorder(img, t, b, l, r, cv2.BORDER_CONSTANT, None, 0))
    same, pad = A.pad_to_shape(img, img.shape)
    assert same is img and pad == (0, 0, 0, 0)


def test_identity_returns_padded_input():
    img = image()
    out = A.transform_img_with_tmat(img, (101, 140), A.IDENTITY.copy())
    assert np.array_equal(out, A.pad_to_shape(img, (101, 140))[0])
    # an identity that takes the general path (not array_equal to the 2x3 identity) is still exact
    out2 = A.warp_fast_bilinear(img, np.eye(3))
    assert np.array_equal(out2.astype(img.dtype), img)


@pytest.mark.parametrize("tx,ty", [(5, 0), (-7, 3), (0, -12), (40, 33)])
def test_integer_translation_is_a_shift(tx, ty):
    img = image()
    t = np.array([[1.0, 0.0, tx], [0.0, 1.0, ty]])
    out = A.transform_img_with_tmat(img, img.shape, t)
    h, w = img.shape
    expect = np.zeros_like(img)
    yy, xx = np.mgrid[0:h, 0:w]
    ok = (yy - ty >= 0) & (yy - ty < h) & (xx - tx >= 0) & (xx - tx < w)
    expect[ok] = img[(yy - ty)[ok], (xx - tx)[ok]]
    # pinv leaves ~1e-16 noise in the matrix: the projective path lands a hair off the integer position and the
    # final truncation may lose one grey level (v * (1 - eps) -> v - 1), never more
    d = expect.astype(np.int64) - out.astype(np.int64)
    assert d.min() >= 0 and d.max() <= 1


def test_dispatch_follows_last_row():
    assert A.transform_kind(np.array([[2.0, 0, 3], [0, 0.5, 1], [0, 0, 1]])) == 0
    assert A.transform_kind(np.array([[2.0, 0.1, 3], [0, 0.5, 1], [0, 0, 1]])) == 1
    assert A.transform_kind(np.array([[2.0, 0.1, 3], [0, 0.5, 1], [1e-18, 0, 1]])) == 2


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("tmat", [
    [[1.0, 0.0, 3.25], [0.0, 1.0, -2.5]],
    [[0.98, -0.05, 4.0], [0.06, 1.01, -3.0]],
    [[1.2, 0.0, -10.0], [0.0, 0.8, 6.0]],
])
def test_against_scipy_affine_transform(dtype, tmat):
    ndi = pytest.importorskip("scipy.ndimage")
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    smooth = cv2.GaussianBlur(rng.random((120, 150)).astype(np.float32), (0, 0), 2)
    img = ((smooth - smooth.min()) / (smooth.max() - smooth.min()) * np.iinfo(dtype).max).astype(dtype)
    tmat = np.array(tmat)
    out = A.transform_img_with_tmat(img, img.shape, tmat)
    M = A.inverse_matrix(tmat)                      # output (x, y) -> input (x, y)
    rc = np.array([[M[1, 1], M[1, 0]], [M[0, 1], M[0, 0]]])   # the same map in (row, col) order
    ref = ndi.affine_transform(img.astype(np.float64), rc, offset=[M[1, 2], M[0, 2]], order=1, mode="constant", cval=0.0)
    # compare where all four neighbours lie inside the input (the two libraries treat the outermost pixel differently)
    yy, xx = np.mgrid[0:img.shape[0], 0:img.shape[1]].astype(np.float64)
    c = M[0, 0] * xx + M[0, 1] * yy + M[0, 2]
    r = M[1, 0] * xx + M[1, 1] * yy + M[1, 2]
    inner = (r > 1) & (r < img.shape[0] - 2) & (c > 1) & (c < img.shape[1] - 2)
    assert inner.mean() > 0.5
    d = np.abs(out.astype(np.float64) - np.floor(ref + 1e-9))[inner]
    assert d.max() <= 1
    outside = (r < -1) | (r > img.shape[0]) | (c < -1) | (c > img.shape[1])
    assert not out[outside].any()


def test_singular_matrix_does_not_crash():
    img = image((40, 50), np.uint8)
    out = A.transform_img_with_tmat(img, img.shape, np.zeros((2, 3)))
    assert out.shape == img.shape and out.dtype == img.dtype
