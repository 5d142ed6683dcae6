"""-m gpu: the alternative window-blur kernels (MA_FB_VARIANT_SHIFT_V / _H, include/microaligner_b200.h) produce the
same stitched flow, bit for bit, as the default kernels.  Only variants already measured on hardware are listed here
(profiles/r01_ab_blur_variants.log); newer ones are exercised by scripts/ab_pipeline.py before they are promoted."""
import numpy as np
import pytest

from tests.util import synth_pair

pytestmark = pytest.mark.gpu

VARIANTS = [(1, 1, 0), (2, 0, 0), (0, 3, 0), (0, 2, 0), (2, 3, 0)]
CASES = {
    "tiled ragged u16": (1300, 1100, np.uint16, 500, 60, 2),
    "untiled odd u8": (333, 415, np.uint8, 0, 0, 2),
    "small window": (500, 460, np.uint16, 200, 8, 1),
}


@pytest.mark.parametrize("case", list(CASES))
def test_blur_variants_bit_identical(cuda, case):
    import torch
    from microaligner_b200 import ops
    h, w, dt, T, ov, it = CASES[case]
    ref, mov = synth_pair(h, w, 11, dt)
    dref, dmov = ops.to_device(ref), ops.to_device(mov)
    win = ov - (1 - ov % 2) if T > 0 else 99
    base = ops.farneback_tiles(dmov, dref, T, ov, win, it, variant=(0, 0, 0), pipelined=False)
    assert torch.isfinite(base).all()
    for v in VARIANTS:
        for contract in (False, True):
            want = base if not contract else ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=True, variant=(0, 0, 0),
                                                                 pipelined=False)
            got = ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=contract, variant=v, pipelined=False)
            assert torch.equal(got, want), f"variant {v} contract_fma={contract}"


def test_unknown_variant_is_rejected(cuda):
    from microaligner_b200 import ops
    from microaligner_b200._lib import MicroalignerB200Error
    ref, mov = synth_pair(300, 300, 1, np.uint16)
    with pytest.raises(MicroalignerB200Error, match="variant"):
        ops.farneback_tiles(ops.to_device(mov), ops.to_device(ref), 0, 0, 99, 1, variant=(9, 0, 0))
