"""-m gpu: the per-cycle pipeline dispatch (z-MIP, serial cycle chain, one flow per cycle applied to all
channels / z-planes, YAML parameter mapping) against the oracle's restatement of __main__.py:320-437."""
import contextlib
import io

import numpy as np
import pytest

from oracle import reference_flow as rf
from tests.util import synth_pair

pytestmark = pytest.mark.gpu

YAML = """
Input:
  ReferenceChannel: DAPI
RegistrationParameters:
  OptFlowReg:
    NumberPyramidLevels: 1
    NumberIterationsPerLevel: 2
    TileSize: 150
    Overlap: 20
    NumberOfWorkers: 0
    UseFullResImage: true
    UseDOG: false
"""


def make_dataset(h=420, w=500, cycles=3, channels=("DAPI", "CD3"), nz=2):
    rng = np.random.default_rng(0)
    ds = {}
    for c in range(1, cycles + 1):
        base, moved = synth_pair(h, w, 10 + c, np.uint16, amp=2.0 + c)
        src = base if c == 1 else moved
        ds[c] = {}
        for ch in channels:
            ds[c][ch] = {}
            for z in range(nz):
                noise = rng.integers(0, 300, (h, w)).astype(np.uint16)
                ds[c][ch][z] = (src // (2 + z) + noise).astype(np.uint16)
    return ds


def test_pipeline_matches_oracle(cuda, tmp_path):
    import yaml
    from microaligner_b200 import pipeline
    cfg = yaml.safe_load(YAML)
    params = pipeline.optflow_parameters(cfg)
    assert params == dict(num_pyr_lvl=1, num_iterations=2, tile_size=150, overlap=20, use_full_res_img=True, use_dog=False)
    ds = make_dataset()
    want = rf.register_cycles(ds, "DAPI", be=rf.CvBackend(), **params)
    got = {}
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        pipeline.run_opt_flow_reg(cfg, ds, lambda cyc, ch, z, img: got.__setitem__((cyc, ch, z), img.copy()))
    assert set(got) == set(want)
    for k in want:
        assert got[k].dtype == want[k].dtype and np.array_equal(got[k], want[k]), k
    out = buf.getvalue()
    assert "Processing Cycle 1 [1/3]" in out and "Skipping as it is a reference image" in out and "Saving Cycle 3 [3/3]" in out
    # YAML file path variant
    p = tmp_path / "config.yaml"
    p.write_text(YAML)
    assert pipeline.optflow_parameters(str(p)) == params


def test_max_projection(cuda):
    from microaligner_b200 import pipeline
    ds = make_dataset(cycles=1, nz=3)
    pages = list(ds[1]["DAPI"].values())
    want = rf.max_project(pages, rf.CvBackend())
    assert np.array_equal(pipeline.max_project_pages(pages).cpu().numpy(), want)
    assert np.array_equal(rf.max_project(pages, rf.NpBackend()), want)


def test_pipeline_from_tiff_files(cuda, tmp_path):
    """C4 in miniature: per-cycle CZYX BigTIFF inputs -> registered (1, C_total, Z, Y, X) BigTIFF stack, pages staged
    through page-locked buffers; compared with the oracle run on the same arrays."""
    import yaml
    from microaligner_b200 import pipeline, tiffio
    cfg = yaml.safe_load(YAML)
    params = pipeline.optflow_parameters(cfg)
    ds = make_dataset(h=380, w=460, cycles=3, channels=("DAPI", "CD3"), nz=2)
    layout, index = {}, {}
    for c, chans in ds.items():
        stack = np.stack([np.stack([chans[ch][z] for z in sorted(chans[ch])]) for ch in chans])   # C, Z, Y, X
        path = tmp_path / f"cycle{c}.tif"
        tiffio.imwrite(path, stack)
        layout[c] = {ch: {z: (str(path), ci * 2 + z) for z in range(2)} for ci, ch in enumerate(chans)}
        for ci, ch in enumerate(chans):
            index[(c, ch)] = (c - 1) * 2 + ci
    prov = tiffio.TiffPageProvider(layout)
    sink = tiffio.TiffStackSink(tmp_path / "registered.tif", index, 2, (380, 460), np.uint16, description="registered")
    with contextlib.redirect_stdout(io.StringIO()):
        pipeline.run_opt_flow_reg(cfg, prov.dataset(), sink)
    sink.close()
    prov.close()
    want = rf.register_cycles(ds, "DAPI", be=rf.CvBackend(), **params)
    with tiffio.TiffFile(tmp_path / "registered.tif") as tif:
        got = tif.asarray().reshape(6, 2, 380, 460)
    for (c, ch), ci in index.items():
        for z in range(2):
            assert np.array_equal(got[ci, z], want[(c, ch, z)]), (c, ch, z)
