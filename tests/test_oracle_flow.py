"""CPU (build container only): the oracle's control-flow restatement against the UNMODIFIED reference
imported from /root/reference through oracle/ref_shim.py.  Skipped where the checkout is absent."""
import contextlib
import io

import numpy as np
import pytest

from oracle import ref_shim
from oracle import reference_flow as rf
from tests.util import synth_pair

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


def run_reference(ref, mov, **kw):
    mod = ref_shim.load()
    r = mod.OptFlowRegistrator()
    for k, v in kw.items():
        setattr(r, k, v)
    r.ref_img, r.mov_img = ref, mov
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        flow = r.register()
    w = mod.Warper()
    w.tile_size, w.overlap = r.tile_size, r.overlap
    w.image, w.flow = mov, flow
    return flow, w.warp(), buf.getvalue()


@pytest.mark.parametrize("case", [
    ((420, 520), np.uint16, dict(num_pyr_lvl=2, tile_size=150, overlap=20, use_full_res_img=True, use_dog=True, num_iterations=2)),
    ((450, 650), np.uint8, dict(num_pyr_lvl=2, tile_size=200, overlap=25, use_full_res_img=False, num_iterations=1)),
])
def test_register_matches_reference(case):
    shape, dtype, kw = case
    ref, mov = synth_pair(shape[0], shape[1], 1, dtype)
    f0, w0, out = run_reference(ref, mov, **kw)
    for be in (rf.CvBackend(), rf.NpBackend()):
        log = []
        f1 = rf.register(ref, mov, be=be, log=log, **kw)
        w1 = rf.warp(mov, f1, kw["tile_size"], kw["overlap"], be)
        assert np.array_equal(f0, f1) and np.array_equal(w0, w1)
        better = [l.strip().startswith("Better") for l in out.splitlines() if "alignment than before" in l]
        assert better == [l["better"] for l in log]


@pytest.mark.parametrize("decisions", [(False, False, False), (True, False, True), (False, True, False), (True, True, False)])
@pytest.mark.parametrize("full_res", [False, True])
def test_forced_decisions_match_reference(monkeypatch, decisions, full_res):
    """The 'Worse alignment' branches (zeros at level 0, x4 at mid levels, x2 / pass-through at the last level,
    optflow_registrator.py:151-169) are rare on real data: force the gate's verdicts in the unmodified reference
    and in the oracle and compare bit for bit."""
    mod = ref_shim.load()
    import importlib
    ofr = importlib.import_module("microaligner.optflow_reg.optflow_registrator")
    ref, mov = synth_pair(420, 500, 3, np.uint16)
    kw = dict(num_pyr_lvl=2, num_iterations=1, tile_size=120, overlap=16, use_full_res_img=full_res)
    dec = list(decisions if full_res else decisions[:2])
    it = iter(dec)
    monkeypatch.setattr(ofr, "check_if_higher_similarity", lambda *a, **k: [next(it)])
    r = mod.OptFlowRegistrator()
    for k, v in kw.items():
        setattr(r, k, v)
    r.ref_img, r.mov_img = ref, mov
    with contextlib.redirect_stdout(io.StringIO()):
        want = r.register()
    got = rf.register(ref, mov, be=rf.CvBackend(), force_decisions=dec, **kw)
    assert want.shape == got.shape and np.array_equal(want, got)
