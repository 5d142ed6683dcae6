"""Dataset structure of a pipeline run: which TIFF page of which file holds z-plane z of channel ch of cycle cyc
(reference microaligner/pipeline_modules/metadata_handling.py:31-158).

The three input layouts differ only in where a cycle's pages start and which file they live in, so one page
enumerator serves all of them.  Keys are 1-based channel and z indices, as in the reference; `ref_channel_ids` holds
the 1-based index of the reference channel inside each cycle."""
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Union

from .ome_meta_processing import (XML, _strip_cycle_info, collect_info_from_ome, generate_ome_for_cycle_builder,
                                  read_ome_meta_from_file)


@dataclass
class DatasetStruct:
    tiff_pages: Dict[int, Dict[int, Dict[int, int]]] = field(default_factory=dict)
    img_paths: Dict[int, Dict[int, Dict[int, Path]]] = field(default_factory=dict)
    ref_channel_ids: Dict[int, int] = field(default_factory=dict)
    ome_xmls: Dict[int, XML] = field(default_factory=dict)

    def add_cycle(self, cyc: int, nchannels: int, nzplanes: int, path_of_channel, first_page: int, pages_per_channel_step: int,
                  ref_channel: int, ome_xml: XML) -> int:
        """Channel ch (1-based) of the cycle is read from path_of_channel(ch); its z-plane z sits at TIFF page
        first_page + (ch - 1) * pages_per_channel_step + (z - 1).  Returns the page after the cycle's last one."""
        pages, paths = {}, {}
        for ch in range(1, nchannels + 1):
            base = first_page + (ch - 1) * pages_per_channel_step
            pages[ch] = {z: base + z - 1 for z in range(1, nzplanes + 1)}
            paths[ch] = {z: path_of_channel(ch) for z in range(1, nzplanes + 1)}
        self.tiff_pages[cyc], self.img_paths[cyc] = pages, paths
        self.ref_channel_ids[cyc] = ref_channel
        self.ome_xmls[cyc] = ome_xml
        return first_page + nchannels * pages_per_channel_step


class DatasetStructCreator:
    def __init__(self):
        self._ref_ch = "DAPI"
        self.img_paths: Union[None, Path, Dict[int, Path], Dict[int, Dict[str, Path]]] = None
        self.input_is_stack = False
        self.input_is_stack_builder = False
        self.output_is_stack = True

    @property
    def ref_channel_name(self) -> str:
        return self._ref_ch

    @ref_channel_name.setter
    def ref_channel_name(self, channel_name: str):
        self._ref_ch = _strip_cycle_info(channel_name)

    def create_dataset_struct(self) -> DatasetStruct:
        if self.img_paths is None:
            raise ValueError("Attribute img_paths is empty")
        ds = DatasetStruct()
        if self.input_is_stack:
            # one file with the channels of all cycles back to back; the distance between the first two occurrences of
            # the reference channel is the number of channels per cycle (metadata_handling.py:97-127)
            path = self.img_paths[sorted(self.img_paths)[0]]
            xml = read_ome_meta_from_file(path)
            info = collect_info_from_ome(self._ref_ch, xml)
            ref_ids = info["ref_ch_ids"]
            per_cycle = ref_ids[1] - ref_ids[0]
            page = 0
            for cyc in range(1, info["nchannels"] // per_cycle + 1):
                page = ds.add_cycle(cyc, per_cycle, info["nzplanes"], lambda ch: path, page, info["nzplanes"], ref_ids[0] + 1, xml)
        elif self.input_is_stack_builder:
            # one single-channel file per channel; pages of a file are its z-planes
            for cyc, xml in generate_ome_for_cycle_builder(self.img_paths).items():
                files = list(self.img_paths[cyc].values())
                info = collect_info_from_ome(self._ref_ch, xml)
                ds.add_cycle(cyc, info["nchannels"], info["nzplanes"], lambda ch, files=files: files[ch - 1], 0, 0,
                             info["ref_ch_ids"][0] + 1, xml)
        else:
            # one multi-channel file per cycle
            for cyc, path in self.img_paths.items():
                xml = read_ome_meta_from_file(path)
                info = collect_info_from_ome(self._ref_ch, xml)
                ds.add_cycle(cyc, info["nchannels"], info["nzplanes"], lambda ch, path=path: path, 0, info["nzplanes"],
                             info["ref_ch_ids"][0] + 1, xml)
        return ds
