"""-m gpu: parity at a size where the production paths are live -- 12 000^2 uint16, default tile geometry (144 full-resolution
tiles): several Farneback workspace batches, the streamed / speculative download of the numpy API, the three-stream host
warp -- against the CPU port of the reference (same cv2 / sklearn calls, all host cores), bit for bit."""
import contextlib
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(900)
def test_12000_squared_bit_identical_to_cpu_port(cuda):
    import cv2
    from benchdata import synth_pair_large
    from microaligner_b200 import OptFlowRegistrator, Warper, ops
    from oracle import reference_flow as rf
    S = 12000
    kw = dict(num_pyr_lvl=3, num_iterations=3, tile_size=1000, overlap=100, use_full_res_img=True, use_dog=False)
    ref, mov = synth_pair_large(S, S, seed=2)
    cv2.setNumThreads(1)
    be = rf.CvBackend(workers=os.cpu_count() or 1)
    log = []
    want_flow = rf.register(ref, mov, be=be, log=log, **kw)
    want_img = rf.warp(mov, want_flow, kw["tile_size"], kw["overlap"], be)
    saved = ops.FARNEBACK_WORKSPACE_BUDGET
    ops.FARNEBACK_WORKSPACE_BUDGET = 6 << 30          # 46 tiles per batch: the full-resolution level runs in batches
    ops.release_workspaces()
    try:
        reg = OptFlowRegistrator()
        for k, v in kw.items():
            setattr(reg, k, v)
        reg.ref_img, reg.mov_img = ref, mov
        with contextlib.redirect_stdout(io.StringIO()):
            flow = reg.register()
        w = Warper()
        w.image, w.flow = mov, flow
        img = w.warp()
    finally:
        ops.FARNEBACK_WORKSPACE_BUDGET = saved
        ops.release_workspaces()
    assert [d["better"] for d in reg.decisions] == [d["better"] for d in log]
    epe = np.sqrt(((flow - want_flow) ** 2).sum(-1))
    assert epe.mean() <= 0.01 and epe.max() <= 0.1      # the written contract (BASELINE.json north_star)
    assert np.array_equal(flow, want_flow), f"flow differs on {np.count_nonzero(flow != want_flow)} values, max EPE {epe.max()}"
    assert np.array_equal(img, want_img)
