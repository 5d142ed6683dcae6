"""`python -m microaligner_b200 config.yaml` -- the reference's command line (`microaligner config.yaml`,
reference microaligner/__main__.py:440-447, 624-642) for the optical-flow half of the pipeline, on the B200.

Same YAML schema (pipeline_modules/config_reader.py), same input layouts (CycleStack, one OME-TIFF per cycle,
CycleBuilder), same output files (`<prefix>optflow_reg_result_stack.tif` or `..._cycNNN.tif`: contiguous BigTIFF with
the reference's OME-XML, written through a memory map), same progress lines on stdout, same per-cycle chain
(__main__.py:320-437): cycle k is registered to the *registered* cycle k-1 and its flow warps every channel and z-plane.

What is different underneath: the z max-projection, the registration and every warp run on the GPU; the flow never
leaves the device; pages are staged through page-locked buffers and uploads / warps / downloads overlap on two streams.
Launched under torchrun (one process per GPU) the registration of each cycle is tile-sharded over all ranks
(engine.py) and the independent channel x z warps of a cycle (__main__.py:427-433) are dealt round-robin to the ranks,
each rank writing its pages straight into the shared output file.

Not available here: the feature-based (affine) registration stage.  A config with a FeatureReg block, or inputs whose
sizes differ (which makes the reference fall back to FeatureReg), is rejected with a clear error."""
import argparse
import os
from pathlib import Path
from pprint import pprint
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops, parallel, tiffio
from .optflow_reg import OptFlowRegistrator, Warper
from .pipeline_modules.config_reader import PipelineConfig, PipelineConfigReader
from .pipeline_modules.metadata_handling import DatasetStruct, DatasetStructCreator
from .pipeline_modules.ome_meta_processing import create_new_meta

Shape2D = Tuple[int, int]
TIMES: Dict[str, float] = {}      # MA_PIPELINE_TIMING=1: device-synchronised wall seconds per stage (scripts/bench_pipeline.py)


class _stage:
    def __init__(self, name):
        self.name, self.on = name, bool(os.environ.get("MA_PIPELINE_TIMING"))

    def __enter__(self):
        if self.on:
            import time
            torch.cuda.synchronize()
            self.t = time.perf_counter()

    def __exit__(self, *a):
        if self.on:
            import time
            torch.cuda.synchronize()
            TIMES[self.name] = TIMES.get(self.name, 0.0) + time.perf_counter() - self.t


def _say(*a):
    if parallel.get().rank == 0:
        print(*a, flush=True)


def _barrier():
    comm = parallel.get()
    if comm.world > 1:
        import torch.distributed as dist
        dist.barrier(group=comm.group)


class PageReader:
    """read_tiff_page (shared_modules/utils.py:69-72) with the files kept open and the pages landing in a small ring
    of page-locked buffers, so the next upload is one DMA transfer."""

    def __init__(self, n_buffers: int = 4):
        self._files: Dict[str, tiffio.TiffFile] = {}
        self._ring: List[np.ndarray] = []
        self._n, self._next = n_buffers, 0
        self.ring_size = n_buffers

    def _page(self, path, index) -> tiffio.TiffPage:
        key = os.fspath(path)
        if key not in self._files:
            self._files[key] = tiffio.TiffFile(key)
        return self._files[key].series[0].pages[index]

    def read(self, path, index, pinned: bool = True) -> np.ndarray:
        page = self._page(path, index)
        dtype = page.dtype.newbyteorder("=")
        if not pinned or not torch.cuda.is_available():
            return page.read_into(np.empty(page.shape, dtype))
        if not self._ring or self._ring[0].shape != page.shape or self._ring[0].dtype != dtype:
            tdt = torch.from_numpy(np.empty(0, dtype)).dtype
            self._ring = [torch.empty(page.shape, dtype=tdt, pin_memory=True).numpy() for _ in range(self._n)]
        buf = self._ring[self._next % self._n]
        self._next += 1
        return page.read_into(buf)

    def close(self):
        for f in self._files.values():
            f.close()
        self._files.clear()


def create_memmap_for_saving(output_path: Path, img_shape, img_dtype, ome_meta: str) -> np.memmap:
    """tif.memmap(..., bigtiff=True, contiguous=True, description=ome_meta) of __main__.py:116-132.  Rank 0 creates the
    file; the other ranks map the same pixel block."""
    comm = parallel.get()
    mm = tiffio.memmap(output_path, img_shape, img_dtype, description=ome_meta) if comm.rank == 0 else None
    _barrier()
    if mm is None:
        mm = tiffio.memmap_existing(output_path, img_shape, img_dtype)
    return mm


def read_and_max_project_pages(reader: PageReader, img_paths: Dict[int, Path], tiff_pages: Dict[int, int]) -> torch.Tensor:
    """z max-projection + 8-bit min-max normalisation (shared_modules/utils.py:75-95) on the device."""
    pages = []
    for z in img_paths:      # page-locked ring buffer -> asynchronous upload; the next page is read meanwhile
        host = reader.read(img_paths[z], tiff_pages[z])
        pages.append(torch.from_numpy(host).to("cuda", non_blocking=True))
        if len(pages) >= reader.ring_size - 1:      # never overwrite a buffer whose upload may still be running
            torch.cuda.current_stream().synchronize()
    return ops.zmip_normalize_u8(pages)


def _my_share(n_items: int) -> List[int]:
    """Indices of the independent pages this rank handles (round-robin over the ranks)."""
    comm = parallel.get()
    return list(range(comm.rank, n_items, comm.world))


def save_pages(mm, reader: PageReader, jobs: List[Tuple[int, int, Path, int]]):
    """First cycle: pages pass through unchanged (__main__.py:305-317).  jobs = (channel index, z index, path, page)."""
    for k in _my_share(len(jobs)):
        ch_id, z_id, path, page = jobs[k]
        tiffio.write_page(mm, (0, ch_id, z_id), reader.read(path, page))


def warp_and_save_pages(mm, reader: PageReader, flow: torch.Tensor, jobs, tile_size: int, overlap: int):
    """warp_and_save_pages (__main__.py:288-302) for this rank's share of a cycle's pages: the flow stays on the device,
    the upload of page k+1 overlaps the warp and the download of page k on two streams."""
    from .engine import Engine
    eng = Engine(tile_size, overlap, comm=parallel.Comm(None))     # whole pages per rank: no sharding inside a page
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    cur = torch.cuda.current_stream()
    pending = []

    def drain(item):
        ch_id, z_id, s, host, _keep = item
        s.synchronize()
        tiffio.write_page(mm, (0, ch_id, z_id), host.numpy())

    for i, k in enumerate(_my_share(len(jobs))):
        ch_id, z_id, path, page = jobs[k]
        s = streams[i % 2]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            img = ops.to_device(reader.read(path, page), flow.device)
            out = eng.warp(img, flow)
            host = torch.empty(out.shape, dtype=out.dtype, device="cpu", pin_memory=True)
            host.copy_(out, non_blocking=True)
        pending.append((ch_id, z_id, s, host, (img, out)))
        if len(pending) > 1:
            drain(pending.pop(0))
    for item in pending:
        drain(item)


def register_and_save_ofreg_imgs(dataset_struct: DatasetStruct, out_dir: Path, filenames: Dict[str, str], tile_size: int,
                                 overlap: int, num_pyr_lvl: int, num_iter: int, ome_meta_per_cyc: Dict[int, str],
                                 input_is_stack: bool, save_to_stack: bool, use_full_res_img: bool, use_dog: bool):
    """Read images and register them sequentially: 1<-2, 2<-3, 3<-4 etc. (__main__.py:320-437, same arguments)."""
    ofreg = OptFlowRegistrator()
    ofreg.tile_size, ofreg.overlap = tile_size, overlap
    ofreg.num_pyr_lvl, ofreg.num_iterations = num_pyr_lvl, num_iter
    ofreg.use_full_res_img, ofreg.use_dog = use_full_res_img, use_dog
    ofreg.gather_flow = True            # page-parallel warps: every rank applies the whole flow to its own pages
    warper = Warper()
    warper.tile_size, warper.overlap = tile_size, overlap

    cycles = list(dataset_struct.tiff_pages.keys())
    first_cycle, ncycles = cycles[0], len(cycles)
    first_channel = next(iter(dataset_struct.img_paths[first_cycle].values()))
    with tiffio.TiffFile(next(iter(first_channel.values()))) as tf:
        img_shape, img_dtype = tf.series[0].shape, tf.series[0].dtype
    max_zplanes = max(len(zs) for cyc in cycles for zs in dataset_struct.tiff_pages[cyc].values())
    nchannels_per_cyc = [len(dataset_struct.tiff_pages[cyc]) for cyc in cycles]
    reader = PageReader()
    decisions = {}

    img_memmap = None
    if save_to_stack:
        shape = (1, sum(nchannels_per_cyc), max_zplanes, img_shape[-2], img_shape[-1])
        img_memmap = create_memmap_for_saving(out_dir / filenames["stack"], shape, img_dtype, ome_meta_per_cyc[first_cycle])

    ref_img = None
    for cyc_id, cyc in enumerate(cycles):
        _say(f"Processing Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
        if not save_to_stack:
            shape = (1, len(dataset_struct.tiff_pages[cyc]), max_zplanes, img_shape[-2], img_shape[-1])
            img_memmap = create_memmap_for_saving(out_dir / filenames["per_cycle"].format(cyc=cyc), shape, img_dtype,
                                                  ome_meta_per_cyc[cyc])
        ref_ch_id = dataset_struct.ref_channel_ids[cyc]
        with _stage("read + max-project"):
            mip = read_and_max_project_pages(reader, dataset_struct.img_paths[cyc][ref_ch_id], dataset_struct.tiff_pages[cyc][ref_ch_id])
        # the independent pages of this cycle: (output channel index, z index, file, TIFF page)
        jobs = []
        for ch_id, ch in enumerate(dataset_struct.tiff_pages[cyc]):
            # in a stack, channel indices run across cycles; per-cycle files restart at 0 ... the reference uses the
            # cross-cycle index for both (__main__.py:413,433), which overruns per-cycle files from the second cycle on;
            # here per-cycle files get the in-file index
            out_ch = cyc_id * nchannels_per_cyc[0] + ch_id if save_to_stack else ch_id
            for z_id, z in enumerate(dataset_struct.img_paths[cyc][ch]):
                jobs.append((out_ch, z_id, dataset_struct.img_paths[cyc][ch][z], dataset_struct.tiff_pages[cyc][ch][z]))
        if cyc == first_cycle:
            _say("Skipping as it is a reference image")
            ref_img = mip
            _say(f"Saving Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
            with _stage("save pages (first cycle)"):
                save_pages(img_memmap, reader, jobs)
        else:
            ofreg.ref_img, ofreg.mov_img = ref_img, mip       # device tensors: the flow stays on the GPU
            with _stage("register"):
                flow = ofreg.register()
            decisions[cyc] = ofreg.decisions
            warper.image, warper.flow = mip, flow
            with _stage("warp of the projection"):
                ref_img = warper.warp()                        # reference of the next cycle
            _say(f"Saving Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
            with _stage("warp + save pages"):
                warp_and_save_pages(img_memmap, reader, flow, jobs, tile_size, overlap)
            del flow
        img_memmap.flush()
        if not save_to_stack:
            _barrier()
            del img_memmap
            img_memmap = None
    if img_memmap is not None:
        _barrier()
        del img_memmap
    reader.close()
    tiffio.close_writers()
    return decisions


def parse_cmd_args(argv=None) -> Path:
    parser = argparse.ArgumentParser(description="MicroAligner: image registration for large scale microscopy")
    parser.add_argument("config", type=Path, help="path to the config yaml file")
    return parser.parse_args(argv).config


def _yx_shapes(img_paths: List[Path]) -> List[Shape2D]:
    shapes = []
    for p in img_paths:
        with tiffio.TiffFile(p) as tf:
            axes, shape = tf.series[0].axes, tf.series[0].shape
            shapes.append((shape[axes.index("Y")], shape[axes.index("X")]))
    return shapes


def get_target_shape(img_paths: List[Path]) -> Shape2D:
    shapes = _yx_shapes(img_paths)
    return max(s[0] for s in shapes), max(s[1] for s in shapes)


def check_input_img_dims_match(img_paths: List[Path]) -> bool:
    shapes = _yx_shapes(img_paths)
    return all(s == shapes[0] for s in shapes)


_NO_FEATURE_REG = ("feature-based (affine) registration is outside the scope of microaligner_b200: run the reference's FeatureReg "
                   "stage first and point InputImagePaths at its *_feature_reg_result_* files")


def run_opt_flow_reg(config: PipelineConfig, img_paths, target_shape: Shape2D):
    """run_opt_flow_reg(config, img_paths, target_shape) of __main__.py:534-609."""
    input_is_stack = config.Input.PipelineInputType == "CycleStack"
    input_is_stack_builder = config.Input.PipelineInputType == "CycleBuilder"
    output_is_stack = config.Output.SaveOutputToCycleStack
    out_dir = Path(config.Output.OutputDir)
    out_prefix = config.Output.OutputPrefix
    p = config.RegistrationParameters.OptFlowReg

    if config.RegistrationParameters.FeatureReg is not None:
        # img_paths are the outputs of a feature-registration stage that ran before (__main__.py:554-556)
        input_is_stack_of = output_is_stack
        input_is_stack_builder = False
    else:
        input_is_stack_of = input_is_stack
        if not input_is_stack_of:
            flat = [q for v in config.Input.InputImagePaths.values() for q in (v.values() if isinstance(v, dict) else [v])]
            if not check_input_img_dims_match([Path(q) for q in flat]):
                raise NotImplementedError("Image dimensions do not match. This probably means that they are not aligned; "
                                          + _NO_FEATURE_REG)
    # NumberOfWorkers configures dask in the reference; here the parallelism comes from the process group (torchrun)

    struct = DatasetStructCreator()
    struct.img_paths = img_paths
    struct.input_is_stack = input_is_stack_of
    struct.input_is_stack_builder = input_is_stack_builder
    struct.output_is_stack = output_is_stack
    struct.ref_channel_name = config.Input.ReferenceChannel
    dataset_struct = struct.create_dataset_struct()

    new_ome_meta = create_new_meta(dataset_struct.ome_xmls, target_shape, input_is_stack_of, output_is_stack)
    names = {"stack": out_prefix + "optflow_reg_result_stack.tif", "per_cycle": out_prefix + "optflow_reg_result_cyc{cyc:03d}.tif"}
    _say("Performing non-linear optical flow based image registration")
    decisions = register_and_save_ofreg_imgs(dataset_struct, out_dir, names, p.TileSize, p.Overlap, p.NumberPyramidLevels,
                                             p.NumberIterationsPerLevel, new_ome_meta, input_is_stack, output_is_stack,
                                             p.UseFullResImage, p.UseDOG)
    _say("Finished\n")
    return decisions


def get_img_path_list(config: PipelineConfig) -> List[Path]:
    paths = config.Input.InputImagePaths
    if config.Input.PipelineInputType == "CycleBuilder":
        return [Path(p) for chans in paths.values() for p in chans.values()]
    return [Path(p) for p in paths.values()]


def _init_ranks():
    """Under torchrun: one process per GPU, NCCL process group, engine sharding installed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and parallel.get().world == 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        parallel.init(dist.group.WORLD)


def main(argv=None):
    _init_ranks()
    _say("Started\n")
    config = PipelineConfigReader().read_config(parse_cmd_args(argv))
    if parallel.get().rank == 0:
        print("The input config is:")
        pprint(config, sort_dicts=False, indent=2)
        config.Output.OutputDir.mkdir(parents=True, exist_ok=True)
    _barrier()
    target_shape = get_target_shape(get_img_path_list(config))
    if config.RegistrationParameters.FeatureReg is not None:
        raise NotImplementedError(_NO_FEATURE_REG)
    if config.RegistrationParameters.OptFlowReg is not None:
        return run_opt_flow_reg(config, config.Input.InputImagePaths, target_shape)


if __name__ == "__main__":
    main()
