"""-m gpu: hypothesis sweeps over ragged shapes and tile geometries, every operator bit-exact vs the oracle."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import cv_ops, farneback_np
from oracle import reference_flow as rf
from tests.util import random_flow

pytestmark = pytest.mark.gpu
SET = dict(deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@settings(max_examples=12, **SET)
@given(h=st.integers(40, 500), w=st.integers(40, 500), T=st.integers(20, 300), u16=st.booleans(), data=st.data())
def test_warp_random_geometry(cuda, h, w, T, u16, data):
    from microaligner_b200 import ops
    ov = data.draw(st.integers(10, max(10, min(T, 50))))
    dtype = np.uint16 if u16 else np.uint8
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, np.iinfo(dtype).max + 1, (h, w)).astype(dtype)
    flow = random_flow(h, w, h + w, mag=5.0)
    got = ops.warp_tiles(dev(img), dev(flow), T, ov).cpu().numpy()
    assert np.array_equal(got, rf.warp(img, flow, T, ov, rf.NpBackend()))


@settings(max_examples=10, **SET)
@given(h=st.integers(30, 400), w=st.integers(30, 400), T=st.integers(20, 200), data=st.data())
def test_merge_random_geometry(cuda, h, w, T, data):
    from microaligner_b200 import ops
    ov = data.draw(st.integers(10, max(10, min(T, 40))))
    f1, f2 = random_flow(h, w, h, 3.0), random_flow(h, w, w, 3.0)
    if data.draw(st.booleans()):
        f1[:T, :T] = 0
    got = ops.merge_flows_tiles(dev(f1), dev(f2), T, ov).cpu().numpy()
    assert np.array_equal(got, rf.merge_flows_tiled(f1, f2, T, ov, rf.NpBackend()))


@settings(max_examples=12, **SET)
@given(h=st.integers(3, 300), w=st.integers(3, 300), u16=st.booleans(), odd_h=st.booleans(), odd_w=st.booleans())
def test_pyramids_random_shapes(cuda, h, w, u16, odd_h, odd_w):
    from microaligner_b200 import ops
    dtype = np.uint16 if u16 else np.uint8
    rng = np.random.default_rng(h * 3 + w)
    img = rng.integers(0, np.iinfo(dtype).max + 1, (h, w)).astype(dtype)
    assert np.array_equal(ops.pyr_down(dev(img)).cpu().numpy(), cv_ops.pyr_down(img))
    f = random_flow(h, w, h + 2 * w)
    dh, dw = 2 * h - int(odd_h), 2 * w - int(odd_w)
    got = ops.pyr_up_flow(dev(f), (dh, dw), 2.0).cpu().numpy()
    assert np.array_equal(got, cv_ops.pyr_up_f32c2(f, (dh, dw), 2.0))


@settings(max_examples=8, **SET)
@given(h=st.integers(24, 260), w=st.integers(24, 260), win=st.sampled_from([9, 19, 39, 99]), iters=st.integers(1, 3),
       u16=st.booleans())
def test_farneback_random_shapes(cuda, h, w, win, iters, u16):
    from microaligner_b200 import ops
    from benchdata import synth_pair
    ref, mov = synth_pair(h, w, h + w, np.uint16 if u16 else np.uint8)
    got = ops.farneback_tiles(dev(mov), dev(ref), 0, 0, win, iters).cpu().numpy()
    assert np.array_equal(got, farneback_np.farneback(mov, ref, win, iters))


@settings(max_examples=8, **SET)
@given(h=st.integers(21, 300), w=st.integers(21, 300), u16=st.booleans())
def test_dog_random_shapes(cuda, h, w, u16):
    from microaligner_b200 import ops
    from benchdata import synth_pair
    ref, _ = synth_pair(h, w, 2 * h + w, np.uint16 if u16 else np.uint8)
    assert np.array_equal(ops.dog_u8(dev(ref)).cpu().numpy(), cv_ops.dog(ref))
