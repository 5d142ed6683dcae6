"""-m gpu: the dependency-cone trimming of the Farneback iterations (FbWin, csrc/farneback.cu) never changes a stitched
value -- the default (trimmed) launch windows against MA_FB_FULL_WINDOWS, with the workspace poisoned in between so
that a value read from outside the trimmed windows would surface as NaN."""
import numpy as np
import pytest

from tests.util import synth_pair

pytestmark = pytest.mark.gpu

CASES = {
    "tiled ragged u16": (1300, 1100, np.uint16, 500, 60, 2),
    "untiled odd u8": (333, 415, np.uint8, 0, 0, 2),
    "small window": (500, 460, np.uint16, 200, 8, 1),
    "default geometry, ragged": (2300, 1500, np.uint16, 1000, 100, 3),
    "deep cone": (900, 700, np.uint8, 300, 150, 4),
    "one iteration": (1100, 1300, np.uint16, 500, 194, 1),
}


@pytest.mark.parametrize("case", list(CASES))
def test_trimmed_iterations_bit_identical(cuda, case):
    import torch
    from microaligner_b200 import ops
    h, w, dt, T, ov, it = CASES[case]
    ref, mov = synth_pair(h, w, 11, dt)
    dref, dmov = ops.to_device(ref), ops.to_device(mov)
    win = ov - (1 - ov % 2) if T > 0 else 99
    for contract in (False, True):
        full = ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=contract, full_windows=True)
        assert torch.isfinite(full).all()
        for ws in ops._FB_WS.values():
            ws.view(torch.float32).fill_(float("nan"))
        got = ops.farneback_tiles(dmov, dref, T, ov, win, it, contract_fma=contract)
        assert torch.equal(got, full), f"trimmed iterations differ from full-window iterations (contract_fma={contract})"


def test_unknown_flag_is_rejected(cuda):
    from microaligner_b200 import _lib, ops
    ref, mov = synth_pair(300, 300, 1, np.uint16)
    r, m = ops.to_device(ref), ops.to_device(mov)
    import torch
    out = torch.empty((300, 300, 2), dtype=torch.float32, device=r.device)
    ws = torch.empty(_lib.lib.ma_farneback_workspace_bytes(300, 300, 0, 0, 1), dtype=torch.uint8, device=r.device)
    rc = _lib.lib.ma_farneback_tiles_ex(m.data_ptr(), r.data_ptr(), 600, _lib.MA_U16, 300, 300, 0, 0, 99, 1, 0, 1, out.data_ptr(),
                                        ws.data_ptr(), ws.numel(), 1 << 9, torch.cuda.current_stream().cuda_stream)
    assert rc == -1 and b"unknown flag" in _lib.lib.ma_last_error()
