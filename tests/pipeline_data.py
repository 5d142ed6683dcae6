"""Seeded miniature datasets for the YAML / TIFF pipeline entry (`python -m microaligner_b200 config.yaml`), shared by
scripts/make_golden.py -- which runs the UNMODIFIED reference on them -- and the tests.

Two of the reference's input layouts (config_reader.py:267-304), both written with the repo's own TIFF writer:

  per_image   one OME-TIFF (C, Z, Y, X) per cycle, channels DAPI + CD3, 2 z-planes      -> one registered stack
  stack       one OME-TIFF holding the channels of all cycles (c01 DAPI, c01 CD3, ...)  -> one registered stack
"""
import os

import numpy as np

from benchdata import synth_pair
from microaligner_b200 import tiffio

H, W = 200, 240
CHANNELS = ("DAPI", "CD3")
PARAMS = dict(NumberPyramidLevels=1, NumberIterationsPerLevel=2, TileSize=100, Overlap=20, NumberOfWorkers=1,
              UseFullResImage=True, UseDOG=False)


def ome_xml(names, nz, h=H, w=W, unit="µm", size=0.5, with_fluor=False):
    ch = "".join(f'<Channel ID="Channel:0:{i}" Name="{n}"' + (f' Fluor="{n}"' if with_fluor else "") + ' SamplesPerPixel="1"/>'
                 for i, n in enumerate(names))
    return ('<?xml version="1.0" encoding="UTF-8"?><OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06" '
            'Creator="tests"><Image ID="Image:0" Name="img"><Pixels ID="Pixels:0" DimensionOrder="XYZCT" Type="uint16" '
            f'SizeX="{w}" SizeY="{h}" SizeZ="{nz}" SizeC="{len(names)}" SizeT="1" PhysicalSizeX="{size}" PhysicalSizeXUnit="{unit}" '
            f'PhysicalSizeY="{size}" PhysicalSizeYUnit="{unit}">{ch}<MetadataOnly/></Pixels></Image></OME>')


def cycle_pages(cyc: int, nz: int):
    """{channel: [z pages]} of one cycle: cycle 1 is the fixed image, later cycles are displaced versions of it."""
    ref, mov = synth_pair(H, W, 40, np.uint16, amp=1.5 + 0.5 * cyc, period=90.0 + 10 * cyc)
    src = ref if cyc == 1 else mov
    rng = np.random.default_rng(100 + cyc)
    out = {}
    for ci, ch in enumerate(CHANNELS):
        planes = []
        for z in range(nz):
            noise = rng.integers(0, 64, (H, W)).astype(np.uint16) * 8
            base = src if ci == 0 else (65535 - src) // 2
            planes.append(((base // (1 + z)) & 0xFFC0) + noise)      # few distinct low bits: the golden file stays small
        out[ch] = planes
    return out


def write_inputs(root, layout: str, ncycles: int = 3):
    """Write the input TIFFs and the YAML config below `root`; returns (config path, output dir)."""
    import yaml
    root = os.fspath(root)
    os.makedirs(root, exist_ok=True)
    out_dir = os.path.join(root, "out")
    if layout == "per_image":
        nz, paths = 2, {}
        for cyc in range(1, ncycles + 1):
            pages = cycle_pages(cyc, nz)
            stack = np.stack([np.stack(pages[ch]) for ch in CHANNELS])          # C, Z, Y, X
            p = os.path.join(root, f"cycle{cyc}.ome.tif")
            tiffio.imwrite(p, stack, description=ome_xml(CHANNELS, nz))
            paths[f"Cycle {cyc}"] = p
    elif layout == "stack":
        nz, names, planes = 1, [], []
        for cyc in range(1, ncycles + 1):
            pages = cycle_pages(cyc, nz)
            for ch in CHANNELS:
                names.append(f"c{cyc:02d} {ch}")
                planes.append(np.stack(pages[ch]))
        p = os.path.join(root, "stack.ome.tif")
        tiffio.imwrite(p, np.stack(planes), description=ome_xml(names, nz, unit="nm", size=325.0))
        paths = {"CycleStack": p}
    else:
        raise ValueError(layout)
    cfg = {"Input": {"InputImagePaths": paths, "ReferenceCycle": 1, "ReferenceChannel": "DAPI"},
           "Output": {"OutputDir": out_dir, "OutputPrefix": "t_", "SaveOutputToCycleStack": True},
           "RegistrationParameters": {"OptFlowReg": dict(PARAMS)}}
    cfg_path = os.path.join(root, "config.yaml")
    with open(cfg_path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return cfg_path, out_dir


def read_result(out_dir):
    """(pixel stack, OME-XML description) of the registered stack."""
    path = os.path.join(os.fspath(out_dir), "t_optflow_reg_result_stack.tif")
    with tiffio.TiffFile(path) as tf:
        return np.array(tf.asarray()), tf.pages[0].description
