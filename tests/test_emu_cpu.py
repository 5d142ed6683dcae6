"""CPU: lane-level emulations of kernel index logic that cannot be run here (no GPU) against the oracle."""
import numpy as np
import pytest

from oracle import farneback_np as fb
from tests import emu_polyexp_march as emu


@pytest.mark.parametrize("shape", [(40, 70), (24, 28), (25, 29), (49, 57), (8, 8), (97, 31), (30, 113)])
def test_marching_polyexp_index_logic(shape):
    """fb_polyexp_march_kernel (csrc/farneback_variants.cuh): virtual rows with REFLECT_101, rolling windows, edge selects,
    shuffle sources and band / strip seams reproduce prefilter3 + polyexp bit for bit."""
    win = np.random.default_rng(shape[0]).integers(0, 65535, shape).astype(np.uint16)
    got = emu.polyexp_march(win)
    assert not np.isnan(got).any()
    assert np.array_equal(got, fb.polyexp(fb.prefilter3(win)))
