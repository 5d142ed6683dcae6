"""CPU: the YAML schema, dataset structure and OME-XML of the pipeline entry (microaligner_b200/pipeline_modules) --
against the golden descriptions produced by the unmodified reference (tests/golden/pipeline_*.npz) and, in the build
container, against the reference's own config reader / dataset-structure code imported through oracle/ref_shim.py."""
import os

import numpy as np
import pytest
import yaml

from microaligner_b200.pipeline_modules import config_reader as cr
from microaligner_b200.pipeline_modules.metadata_handling import DatasetStructCreator
from microaligner_b200.pipeline_modules.ome_meta_processing import (_strip_cycle_info, collect_info_from_ome, create_new_meta,
                                                                   read_ome_meta_from_file, str_to_xml)
from oracle import ref_shim
from tests import pipeline_data as pd

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _struct(config):
    s = DatasetStructCreator()
    s.img_paths = config.Input.InputImagePaths
    s.input_is_stack = config.Input.PipelineInputType == "CycleStack"
    s.input_is_stack_builder = config.Input.PipelineInputType == "CycleBuilder"
    s.output_is_stack = config.Output.SaveOutputToCycleStack
    s.ref_channel_name = config.Input.ReferenceChannel
    return s.create_dataset_struct()


@pytest.mark.parametrize("layout", ["per_image", "stack"])
def test_description_and_config_text_match_the_reference(tmp_path, layout):
    cfg_path, _ = pd.write_inputs(tmp_path, layout)
    gold = np.load(os.path.join(GOLDEN, f"pipeline_{layout}.npz"))
    config = cr.PipelineConfigReader().read_config(cfg_path)
    # the reference pretty-prints the config object; the text must be the same
    import io
    from pprint import pprint
    buf = io.StringIO()
    pprint(config, stream=buf, sort_dicts=False, indent=2)
    assert buf.getvalue().replace(str(tmp_path), "<TMP>").strip() in str(gold["stdout"])
    ds = _struct(config)
    meta = create_new_meta(ds.ome_xmls, (pd.H, pd.W), config.Input.PipelineInputType == "CycleStack", True)
    assert meta[list(meta)[0]] == str(gold["description"])
    assert list(ds.tiff_pages) == [1, 2, 3] and all(ds.ref_channel_ids[c] == 1 for c in ds.tiff_pages)
    if layout == "per_image":
        assert ds.tiff_pages[2] == {1: {1: 0, 2: 1}, 2: {1: 2, 2: 3}}
    else:
        assert ds.tiff_pages[2] == {1: {1: 2}, 2: {1: 3}} and ds.tiff_pages[3] == {1: {1: 4}, 2: {1: 5}}


def test_output_layouts_and_units(tmp_path):
    cfg_path, _ = pd.write_inputs(tmp_path, "per_image")
    ds = _struct(cr.PipelineConfigReader().read_config(cfg_path))
    per_file = create_new_meta(ds.ome_xmls, (pd.H, pd.W), False, False)
    assert set(per_file) == {1, 2, 3}
    x = str_to_xml(per_file[2])
    px = x.find("Image").find("Pixels")
    assert px.get("SizeC") == "2" and px.get("PhysicalSizeX") == "500.0" and px.get("PhysicalSizeXUnit") == "nm"
    assert len(px.findall("TiffData")) == 4 and [c.get("Name") for c in px.findall("Channel")] == ["DAPI", "CD3"]
    cfg_path, _ = pd.write_inputs(tmp_path / "s", "stack")
    ds = _struct(cr.PipelineConfigReader().read_config(cfg_path))
    split = create_new_meta(ds.ome_xmls, (pd.H, pd.W), True, False)
    names = [[c.get("Name") for c in str_to_xml(split[c]).find("Image").find("Pixels").findall("Channel")] for c in (1, 2, 3)]
    assert names == [["c01 DAPI", "c01 CD3"], ["c02 DAPI", "c02 CD3"], ["c03 DAPI", "c03 CD3"]]
    assert _strip_cycle_info("cyc12_DAPI-3") == "DAPI" and _strip_cycle_info("c01 CD3") == "CD3"
    with pytest.raises(ValueError, match="Incorrect reference channel"):
        collect_info_from_ome("XYZ", read_ome_meta_from_file(ds.img_paths[1][1][1]))


BAD = [
    ({"Input": 1}, ValueError, "absent"),
    ("drop:Input.ReferenceCycle", KeyError, "Field ReferenceCycle is absent"),
    ("set:Input.ReferenceCycle=0", ValueError, "smaller than minimum: 1"),
    ("set:Input.ReferenceChannel=3", TypeError, "Field ReferenceChannel has wrong data type"),
    ("set:RegistrationParameters.OptFlowReg.Overlap=150", ValueError, "Field Overlap value is greater than maximum: 100"),
    ("set:RegistrationParameters.OptFlowReg.TileSize=10", ValueError, "Field TileSize value is smaller than minimum: 20"),
    ("set:RegistrationParameters.OptFlowReg.NumberPyramidLevels=9", ValueError, "greater than maximum: 8"),
    ("set:RegistrationParameters.OptFlowReg.UseDOG=1", TypeError, "Field UseDOG has wrong data type"),
    ("drop:RegistrationParameters.OptFlowReg", ValueError, "At least one of the registration methods"),
    ("set:Input.InputImagePaths={'Cycle 1': 'a.tif'}", ValueError, "Not enough cycles"),
    ("set:Input.InputImagePaths={'cycle1': 'a.tif', 'cycle2': 'b.tif'}", ValueError, "should follow pattern Cycle N"),
    ("set:Input.InputImagePaths={'CycleStack': 'a.tif', 'Cycle 2': 'b.tif'}", ValueError, "at most 1 image path"),
]


def _mutate(cfg, spec):
    if isinstance(spec, dict):
        return spec
    op, rest = spec.split(":", 1)
    path, _, value = rest.partition("=")
    keys = path.split(".")
    node = cfg
    for k in keys[:-1]:
        node = node[k]
    if op == "drop":
        del node[keys[-1]]
    else:
        node[keys[-1]] = eval(value)      # noqa: S307 -- literals of this file
    return cfg


@pytest.mark.parametrize("spec,exc,msg", BAD)
def test_bad_configs_fail_like_the_reference(tmp_path, spec, exc, msg):
    cfg_path, _ = pd.write_inputs(tmp_path, "per_image")
    cfg = _mutate(yaml.safe_load(open(cfg_path)), spec)
    bad = tmp_path / "bad.yaml"
    bad.write_text(yaml.safe_dump(cfg, sort_keys=False))
    with pytest.raises(exc, match=msg) as mine:
        cr.PipelineConfigReader().read_config(bad)
    if ref_shim.available():       # same exception type and text from the unmodified reference
        ref_shim.load_pipeline()
        import importlib
        ref_cr = importlib.import_module("microaligner.pipeline_modules.config_reader")
        with pytest.raises(exc) as theirs:
            ref_cr.PipelineConfigReader().read_config(bad)
        assert str(theirs.value) == str(mine.value)


def test_cycle_builder_layout(tmp_path):
    """CycleBuilder input: one single-channel file per channel; the synthetic OME-XML names the channels after the keys."""
    from microaligner_b200 import tiffio
    paths = {}
    for cyc in (1, 2):
        pages = pd.cycle_pages(cyc, 1)
        paths[f"Cycle {cyc}"] = {}
        for ch in pd.CHANNELS:
            p = tmp_path / f"c{cyc}_{ch}.tif"
            tiffio.imwrite(p, pages[ch][0])
            paths[f"Cycle {cyc}"][ch] = str(p)
    cfg = {"Input": {"InputImagePaths": paths, "ReferenceCycle": 1, "ReferenceChannel": "CD3"},
           "Output": {"OutputDir": str(tmp_path / "o"), "OutputPrefix": "", "SaveOutputToCycleStack": True},
           "RegistrationParameters": {"OptFlowReg": dict(pd.PARAMS)}}
    f = tmp_path / "cb.yaml"
    f.write_text(yaml.safe_dump(cfg, sort_keys=False))
    config = cr.PipelineConfigReader().read_config(f)
    assert config.Input.PipelineInputType == "CycleBuilder"
    ds = _struct(config)
    assert ds.ref_channel_ids == {1: 2, 2: 2} and ds.tiff_pages[1] == {1: {1: 0}, 2: {1: 0}}
    assert [str(ds.img_paths[2][c][1]) for c in (1, 2)] == [paths["Cycle 2"][ch] for ch in pd.CHANNELS]
    if ref_shim.available():
        ref_shim.load_pipeline()
        import importlib
        ref_mh = importlib.import_module("microaligner.pipeline_modules.metadata_handling")
        ref_cr = importlib.import_module("microaligner.pipeline_modules.config_reader")
        rc = ref_cr.PipelineConfigReader().read_config(f)
        s = ref_mh.DatasetStructCreator()
        s.img_paths, s.input_is_stack, s.input_is_stack_builder, s.output_is_stack = rc.Input.InputImagePaths, False, True, True
        s.ref_channel_name = "CD3"
        theirs = s.create_dataset_struct()
        assert theirs.tiff_pages == ds.tiff_pages and theirs.img_paths == ds.img_paths and theirs.ref_channel_ids == ds.ref_channel_ids
        ome = importlib.import_module("microaligner.pipeline_modules.ome_meta_processing")
        assert ome.create_new_meta(theirs.ome_xmls, (pd.H, pd.W), False, True) == create_new_meta(ds.ome_xmls, (pd.H, pd.W), False, True)
