"""-m gpu: ma_warp_affine / transform_img_with_tmat / pad_to_shape against oracle/affine_np.py
(reference shared_modules/utils.py:40-66, 98-114; parity unpinned against scikit-image itself, see the oracle header).
Bit-exact: the kernel and the oracle perform the same float64 operations in the same order."""
import numpy as np
import pytest

from oracle import affine_np as A
from tests.util import blobs

pytestmark = pytest.mark.gpu

TMATS = {
    "translation": [[1.0, 0.0, 3.25], [0.0, 1.0, -2.5]],            # pinv noise -> projective path
    "integer shift": [[1.0, 0.0, 12.0], [0.0, 1.0, -7.0]],
    "rotation+scale": [[0.98, -0.05, 4.0], [0.06, 1.01, -3.0]],
    "anisotropic": [[1.2, 0.0, -10.0], [0.0, 0.8, 6.0]],
    "shear": [[1.0, 0.3, 0.0], [0.0, 1.0, 0.0]],
    "singular": [[0.0, 0.0, 5.0], [0.0, 0.0, 5.0]],
    "far away": [[1.0, 0.0, 1e7], [0.0, 1.0, 0.0]],
}


def page(shape, dtype, seed=0):
    if dtype == np.uint16:
        return blobs(shape[0], shape[1], seed, np.uint16)
    rng = np.random.default_rng(seed)
    return rng.integers(0, 255, shape, endpoint=True).astype(np.uint8)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16])
@pytest.mark.parametrize("name", list(TMATS))
def test_transform_img_with_tmat(cuda, dtype, name):
    from microaligner_b200 import transform_img_with_tmat
    img = page((301, 413), dtype, 3)
    tmat = np.array(TMATS[name])
    for target in [img.shape, (330, 450), (302, 416)]:
        want = A.transform_img_with_tmat(img, target, tmat)
        got = transform_img_with_tmat(img, target, tmat)
        assert got.dtype == img.dtype and got.shape == tuple(target)
        assert np.array_equal(got, want), f"{name} {target}: {np.count_nonzero(got != want)} px differ"


@pytest.mark.parametrize("kind,M", [
    (0, [[1.25, 0.0, -3.5], [0.0, 0.75, 2.25], [0.0, 0.0, 1.0]]),
    (1, [[1.0, 0.125, -3.5], [-0.0625, 1.0, 2.25], [0.0, 0.0, 1.0]]),
    (2, [[1.0, 0.125, -3.5], [-0.0625, 1.0, 2.25], [1e-5, -2e-5, 1.0]]),
])
def test_three_coordinate_paths(cuda, kind, M):
    """metric / affine / projective dispatch of _warp_fast, driven directly through the C ABI wrapper."""
    from microaligner_b200 import ops
    M = np.array(M)
    assert A.transform_kind(M) == kind
    img = page((200, 260), np.uint16, 5)
    want = A.warp_fast_bilinear(img, M)
    want = np.clip(want, 0, img.max()).astype(np.uint16)
    got = ops.to_host(ops.warp_affine(ops.to_device(img), M, img.shape))
    assert np.array_equal(got, want)


def test_pad_to_shape_and_identity(cuda):
    import torch
    from microaligner_b200 import pad_to_shape, transform_img_with_tmat
    img = page((120, 97), np.uint16, 1)
    for target in [(120, 97), (121, 97), (127, 104), (130, 98)]:
        want, wpad = A.pad_to_shape(img, target)
        got, pad = pad_to_shape(img, target)
        assert pad == wpad and np.array_equal(got, want)
        ident = transform_img_with_tmat(img, target, np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]))
        assert np.array_equal(ident, want)
    same, pad = pad_to_shape(img, img.shape)
    assert same is img and pad == (0, 0, 0, 0)
    dev = torch.from_numpy(img).cuda()
    out, _ = pad_to_shape(dev, (127, 104))
    assert isinstance(out, torch.Tensor) and out.is_cuda
    assert np.array_equal(out.cpu().numpy(), A.pad_to_shape(img, (127, 104))[0])
    with pytest.raises(ValueError):
        pad_to_shape(img, (100, 97))


def test_transform_and_save_zplanes(cuda):
    from microaligner_b200 import pipeline
    tmat = np.array(TMATS["rotation+scale"])
    pages = {z: page((150, 170), np.uint16, 10 + z) for z in range(3)}
    got = {}
    pipeline.transform_and_save_zplanes(lambda cyc, ch, z, im: got.__setitem__((cyc, ch, z), im.copy()), 2, "CD3", (160, 176),
                                        tmat, pages, max_zplanes=5)
    assert sorted(got) == [(2, "CD3", z) for z in range(5)]
    for z in range(3):
        assert np.array_equal(got[(2, "CD3", z)], A.transform_img_with_tmat(pages[z], (160, 176), tmat))
    for z in (3, 4):
        assert got[(2, "CD3", z)].shape == (160, 176) and not got[(2, "CD3", z)].any()
