"""OME-XML bookkeeping of the pipeline (reference microaligner/pipeline_modules/ome_meta_processing.py and
stack_builder.py): read the description of the input TIFFs, find the reference channel, and write the description
of the output files for the four input / output layouts (stack or one file per cycle on either side).

Nothing here touches pixels.  The strings produced are byte-identical to the reference's for the same inputs (same
element and attribute order through xml.etree), except that physical pixel sizes are converted to nanometres with
an exact table of SI prefixes instead of the `pint` package, which is not part of this environment."""
import re
import xml.etree.ElementTree as ET
from copy import deepcopy
from io import StringIO
from pathlib import Path
from typing import Any, Dict, List, Sequence, Tuple

from .. import tiffio

XML = ET.Element
OME_ATTRIBS = {
    "xmlns": "http://www.openmicroscopy.org/Schemas/OME/2016-06",
    "xmlns:xsi": "http://www.w3.org/2001/XMLSchema-instance",
    # the blank inside "openmicr oscopy" is what the reference writes (ome_meta_processing.py:267); kept for identical files
    "xsi:schemaLocation": "http://www.openmicr oscopy.org/Schemas/OME/2016-06 http://www.openmicroscopy.org/Schemas/OME/2016-06/ome.xsd",
}
XML_DECLARATION = '<?xml version="1.0" encoding="UTF-8"?>'
# length unit -> nanometres (what pint's registry resolves these symbols to)
_NM_PER_UNIT = {"nm": 1.0, "um": 1e3, "µm": 1e3, "μm": 1e3, "micron": 1e3, "mm": 1e6, "cm": 1e7, "m": 1e9, "pm": 1e-3,
                "angstrom": 0.1, "Å": 0.1, "inch": 25.4e6, "in": 25.4e6}


def str_to_xml(text: str) -> XML:
    """Parse and drop the namespace prefix of every tag."""
    it = ET.iterparse(StringIO(text))
    for _, el in it:
        el.tag = el.tag.rpartition("}")[2]
    return it.root


def xml_to_string(xml: XML) -> str:
    return XML_DECLARATION + ET.tostring(xml, method="xml", encoding="utf-8").decode("ascii", errors="ignore")


def read_ome_meta_from_file(path) -> XML:
    with tiffio.TiffFile(Path(path)) as tf:
        text = tf.ome_metadata
    if not text:
        raise ValueError(f"{path} carries no OME-XML description")
    return str_to_xml(text)


def _pixels(xml: XML) -> XML:
    return xml.find("Image").find("Pixels")


def _strip_cycle_info(name: str) -> str:
    """'c01 DAPI', 'cyc2_DAPI-1' -> 'DAPI' (ome_meta_processing.py:69-72)."""
    name = re.sub(r"^(c|cyc|cycle)\d+(\s+|_|-)?", "", name)
    return re.sub(r"(-\d+)?(_\d+)?$", "", name)


def to_nm(value: float, unit: str) -> float:
    try:
        return value * _NM_PER_UNIT[unit]
    except KeyError:
        raise ValueError(f"unknown length unit {unit!r} in the OME-XML PhysicalSize attributes") from None


def channel_info(xml: XML) -> Dict[str, Any]:
    px = _pixels(xml)
    channels = px.findall("Channel")
    return {
        "channels": channels,
        "channel_names": [c.get("Name") for c in channels],
        "channel_fluors": [c.get("Fluor") for c in channels if "Fluor" in c.attrib],
        "nchannels": int(px.attrib.get("SizeC", 1)),
        "nzplanes": int(px.attrib.get("SizeZ", 1)),
    }


def pixels_info(xml: XML) -> Dict[str, Any]:
    px = _pixels(xml)
    info: Dict[str, Any] = {d: int(px.get(d, 1)) for d in ("SizeX", "SizeY", "SizeC", "SizeZ", "SizeT")}
    info.update({s: float(px.get(s, 1)) for s in ("PhysicalSizeX", "PhysicalSizeY")})
    info.update({u: px.get(u, "um") for u in ("PhysicalSizeXUnit", "PhysicalSizeYUnit")})
    return info


def collect_info_from_ome(ref_ch: str, xml: XML) -> Dict[str, Any]:
    """Channel / size information of one file plus `ref_ch_ids`: the indices of the channels whose cleaned name (or
    fluorophore) starts with the reference channel's name, case-insensitively (re.match, as in the reference)."""
    info = channel_info(xml)
    names = [_strip_cycle_info(n) for n in info["channel_names"]]
    fluors = [_strip_cycle_info(f) for f in info["channel_fluors"]] or None
    if ref_ch in names:
        cleaned = names
    elif fluors is not None and ref_ch in fluors:
        cleaned = fluors
    else:
        extra = f", fluors: {set(fluors)}" if fluors is not None else ""
        raise ValueError(f"Incorrect reference channel {ref_ch}. Available channel names: {set(names)}{extra}")
    out = dict(info)
    out["ref_ch_ids"] = [i for i, ch in enumerate(cleaned) if re.match(ref_ch, ch, re.IGNORECASE)]
    out.update(pixels_info(xml))
    return out


# --------------------------------------------------------------------------------------- output descriptions
def _sizes(xmls: Sequence[XML], target_shape: Tuple[int, int]) -> Dict[str, Any]:
    """Pixels attributes of an output holding the channels of all `xmls`, in nanometres."""
    infos = [pixels_info(x) for x in xmls]
    return {
        "SizeX": target_shape[1], "SizeY": target_shape[0],
        "SizeC": sum(i["SizeC"] for i in infos), "SizeZ": max(i["SizeZ"] for i in infos), "SizeT": max(i["SizeT"] for i in infos),
        "PhysicalSizeX": to_nm(max(i["PhysicalSizeX"] for i in infos), infos[-1]["PhysicalSizeXUnit"]),
        "PhysicalSizeY": to_nm(max(i["PhysicalSizeY"] for i in infos), infos[-1]["PhysicalSizeYUnit"]),
        "PhysicalSizeXUnit": "nm", "PhysicalSizeYUnit": "nm",
    }


def _tiff_data(n_t: int, n_c: int, n_z: int) -> List[XML]:
    nodes, ifd = [], 0
    for t in range(n_t):
        for c in range(n_c):
            for z in range(n_z):
                nodes.append(ET.Element("TiffData", {"FirstC": str(c), "FirstT": str(t), "FirstZ": str(z), "IFD": str(ifd),
                                                     "PlaneCount": "1"}))
                ifd += 1
    return nodes


def _rewrite(template: XML, sizes: Dict[str, Any], channels=None) -> str:
    """Copy of `template` with the canonical OME root attributes, the output's sizes, optionally a new channel list
    and fresh TiffData nodes (appended after whatever other children Pixels keeps)."""
    xml = deepcopy(template)
    xml.attrib.clear()
    for k, v in OME_ATTRIBS.items():
        xml.set(k, v)
    px = _pixels(xml)
    px.set("DimensionOrder", "XYZCT")
    for k, v in sizes.items():
        px.set(k, str(v))
    for old in px.findall("TiffData") + (px.findall("Channel") if channels is not None else []):
        px.remove(old)
    for node in list(channels or []) + _tiff_data(sizes["SizeT"], sizes["SizeC"], sizes["SizeZ"]):
        px.append(node)
    return xml_to_string(xml)


def _renumbered(channels: Sequence[XML], names: Sequence[str], first_id: int = 0) -> List[XML]:
    out = []
    for i, (ch, name) in enumerate(zip(channels, names)):
        ch = deepcopy(ch)
        ch.set("Name", name)
        ch.set("ID", "Channel:0:" + str(first_id + i))
        out.append(ch)
    return out


def create_new_meta(ome_meta_per_cyc: Dict[int, XML], target_shape: Tuple[int, int], input_is_stack: bool,
                    output_is_stack: bool) -> Dict[int, str]:
    """{cycle: OME-XML string of the file that cycle is written to} (ome_meta_processing.py:454-473)."""
    cycles = list(ome_meta_per_cyc)
    xmls = [ome_meta_per_cyc[c] for c in cycles]
    if input_is_stack and output_is_stack:          # stack in, stack out: the description passes through
        return {c: xml_to_string(x) for c, x in ome_meta_per_cyc.items()}
    if output_is_stack:                             # files in, one stack out: channels of all cycles, prefixed "cNN "
        width = len(str(len(cycles))) + 1
        channels: List[XML] = []
        for i, x in enumerate(xmls):
            info = channel_info(x)
            names = ["c" + format(i + 1, f"0{width}d") + " " + n for n in info["channel_names"]]
            channels += _renumbered(info["channels"], names, len(channels))
        text = _rewrite(xmls[0], _sizes(xmls, target_shape), channels)
        return {c: text for c in cycles}
    if input_is_stack:                              # stack in, one file per cycle out: slice the channel list
        per_cyc = int(round(pixels_info(xmls[0])["SizeC"] / len(cycles), 0))
        out = {}
        for n, (c, x) in enumerate(zip(cycles, xmls)):
            sizes = _sizes([x], target_shape)
            sizes["SizeC"] = per_cyc
            info = channel_info(x)
            sl = slice(n * per_cyc, (n + 1) * per_cyc)
            out[c] = _rewrite(x, sizes, _renumbered(info["channels"][sl], info["channel_names"][sl]))
        return out
    # files in, files out
    return {c: _rewrite(x, _sizes([x], target_shape), None) for c, x in zip(cycles, xmls)}


# --------------------------------------------------------------------------------------- CycleBuilder input
def _plain_image_dims(path) -> Dict[str, int]:
    """Y, X and the number of pages of a plain (non-OME) single-channel TIFF used as CycleBuilder input."""
    with tiffio.TiffFile(Path(path)) as tf:
        s = tf.series[0]
        dims = dict(zip(s.axes, s.shape))
    higher = [v for k, v in dims.items() if k not in ("Y", "X") and v > 1]
    if len(higher) >= 2:
        raise ValueError("The input image has too many dimensions")
    return {"Y": dims["Y"], "X": dims["X"], "Z": higher[0] if higher else 1}


def generate_ome_for_cycle_builder(cycle_map: Dict[int, Dict[str, Path]]) -> Dict[int, XML]:
    """One synthetic OME-XML tree per cycle for inputs given as {cycle: {channel name: file}} (stack_builder.py:216-227)."""
    first_path = next(iter(next(iter(cycle_map.values())).values()))
    with tiffio.TiffFile(Path(first_path)) as tf:
        pixels_attrib = {"ID": "Pixels:0", "DimensionOrder": "XYZCT", "Interleaved": "false", "Type": tf.series[0].dtype.name}
    root_attrib = dict(OME_ATTRIBS)
    root_attrib["xsi:schemaLocation"] = root_attrib["xsi:schemaLocation"].replace("openmicr oscopy", "openmicroscopy")
    out, offset = {}, 0
    for cyc, channels in cycle_map.items():
        names = list(channels)
        d = _plain_image_dims(channels[names[0]])
        dims = {"SizeT": 1, "SizeZ": 1 if d["Z"] == 1 else d["Z"] * len(names), "SizeC": len(names), "SizeY": d["Y"], "SizeX": d["X"]}
        pixels_attrib.update({k: str(v) for k, v in dims.items()})
        ome = ET.Element("OME", root_attrib)
        image = ET.SubElement(ome, "Image", {"ID": "Image:0", "Name": "default.tif"})
        px = ET.SubElement(image, "Pixels", pixels_attrib)
        for i, name in enumerate(names):
            ET.SubElement(px, "Channel", {"ID": "Channel:0:" + str(offset + i), "Name": name, "SamplesPerPixel": "1"})
        ifd = 0
        for t in range(dims["SizeT"]):
            for c in range(dims["SizeC"]):
                for z in range(dims["SizeZ"]):
                    ET.SubElement(px, "TiffData", {"FirstT": str(t), "FirstC": str(c), "FirstZ": str(z), "IFD": str(ifd)})
                    ifd += 1
        offset += len(names)
        out[cyc] = str_to_xml(XML_DECLARATION + ET.tostring(ome, encoding="utf-8", method="xml").decode("ascii"))
    return out
