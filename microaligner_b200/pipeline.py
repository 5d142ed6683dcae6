"""Per-cycle dispatch of the optical-flow registration pipeline on the B200
(reference microaligner/__main__.py: register_and_save_ofreg_imgs :320-437, warp_and_save_pages :288-302,
run_opt_flow_reg :534-609, read_and_max_project_pages shared_modules/utils.py:75-95).

What stays the same: the YAML keys (RegistrationParameters.OptFlowReg.*), the serial chain over cycles
(cycle k is registered to the *registered* cycle k-1), the z max-projection + 8-bit normalisation that
feeds register(), one flow per cycle applied to every channel and z-plane of that cycle, the progress
lines on stdout.

What changes: the flow never leaves the GPU (register() hands a device tensor to the warps), the
max-projection runs on the device, and page uploads / warps / downloads are double-buffered on two CUDA
streams.  TIFF / OME-XML handling is out of scope (tifffile is not part of this environment): pages come
from a *page provider* -- any mapping  cycle -> channel -> z -> 2-D uint8/uint16 array (numpy or CUDA
tensor; values may be callables returning the array for lazy loading) -- and results go to a *sink*
callable(cycle, channel, z_index, image: np.ndarray), z_index = position of the plane among the channel's planes
(the reference's datasets use 1-based z keys but write plane z_id = 0, 1, ...: __main__.py:296-300).
The file-based entry point with the reference's YAML / TIFF handling is microaligner_b200/__main__.py."""
from typing import Callable, Dict, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import ops, parallel
from .engine import Engine
from .optflow_reg import OptFlowRegistrator, Warper
from .shared_modules.utils import transform_img_with_tmat

Page = Union[np.ndarray, torch.Tensor, Callable[[], np.ndarray]]
Dataset = Mapping[int, Mapping[str, Mapping[int, Page]]]


def _load(page: Page):
    return page() if callable(page) else page


def max_project_pages(pages: Sequence[Page]) -> torch.Tensor:
    """read_and_max_project_pages (utils.py:75-95): np.maximum over the z-planes, then
    cv.normalize(.., 0, 255, NORM_MINMAX, CV_8U) -- on the device, one uint8 CUDA tensor out."""
    dev_pages = [ops.to_device(_load(p)) for p in pages]
    return ops.zmip_normalize_u8(dev_pages)


def optflow_parameters(config: Union[str, Mapping]) -> Dict:
    """RegistrationParameters.OptFlowReg block of a pipeline YAML (path or parsed dict) -> attribute values,
    the mapping of __main__.py:543-550.  NumberOfWorkers (a dask setting there) is accepted and ignored:
    the device path takes its parallelism from the process group (parallel.init)."""
    if isinstance(config, str):
        import yaml
        with open(config) as f:
            config = yaml.safe_load(f)
    p = config["RegistrationParameters"]["OptFlowReg"]
    return dict(num_pyr_lvl=int(p["NumberPyramidLevels"]), num_iterations=int(p["NumberIterationsPerLevel"]),
                tile_size=int(p["TileSize"]), overlap=int(p["Overlap"]),
                use_full_res_img=bool(p.get("UseFullResImage", False)), use_dog=bool(p.get("UseDOG", False)))


def _say(*a):
    if parallel.get().rank == 0:
        print(*a)


def _rank0_sink(sink):
    """With several ranks (parallel.init) every rank runs the same SPMD loop; only rank 0 hands results to the sink."""
    return sink if parallel.get().rank == 0 else (lambda *a, **k: None)


def warp_and_save_pages(sink, cyc, ch, flow: torch.Tensor, pages: Mapping[int, Page], tile_size: int, overlap: int):
    """warp_and_save_pages (__main__.py:288-302) with the flow resident on the device; uploads, warps and
    downloads of consecutive z-planes overlap on two streams."""
    sink = _rank0_sink(sink)
    eng = Engine(tile_size, overlap, comm=parallel.get())
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    cur = torch.cuda.current_stream()
    pending = []
    for z_id, page in enumerate(pages.values()):      # the sink gets the plane INDEX, like mm[0, ch, z_id] (__main__.py:296-300)
        s = streams[z_id % 2]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            img = ops.to_device(_load(page), flow.device)
            out = eng.warp(img, flow)
            host = torch.empty(out.shape, dtype=out.dtype, device="cpu", pin_memory=True)
            host.copy_(out, non_blocking=True)
        pending.append((z_id, s, host, img, out))
        if len(pending) > 1:
            z0, s0, h0, _, _ = pending.pop(0)
            s0.synchronize()
            sink(cyc, ch, z0, h0.numpy())
    for z0, s0, h0, _, _ in pending:
        s0.synchronize()
        sink(cyc, ch, z0, h0.numpy())


def transform_and_save_zplanes(sink, cyc, ch, target_shape, transform_matrix, pages: Mapping[int, Page], max_zplanes: int):
    """transform_and_save_zplanes (__main__.py:84-112): every z-plane of one channel is padded to target_shape and
    resampled by the cycle's 2x3 matrix (transform_img_with_tmat); channels with fewer planes than max_zplanes are
    completed with empty pages.  Uploads, resampling and downloads of consecutive planes overlap on two streams."""
    sink = _rank0_sink(sink)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    cur = torch.cuda.current_stream()
    pending, z_id, last = [], 0, None
    for i, (z, page) in enumerate(pages.items()):
        s = streams[i % 2]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            img = ops.to_device(_load(page))
            out = transform_img_with_tmat(img, target_shape, transform_matrix)
            host = torch.empty(out.shape, dtype=out.dtype, device="cpu", pin_memory=True)
            host.copy_(out, non_blocking=True)
        pending.append((z_id, s, host, img, out))
        z_id += 1
        if len(pending) > 1:
            z0, s0, h0, _, _ = pending.pop(0)
            s0.synchronize()
            sink(cyc, ch, z0, h0.numpy())
            last = h0
    for z0, s0, h0, _, _ in pending:
        s0.synchronize()
        sink(cyc, ch, z0, h0.numpy())
        last = h0
    if last is not None:
        for z0 in range(z_id, max_zplanes):
            sink(cyc, ch, z0, np.zeros_like(last.numpy()))


def register_and_save_ofreg_imgs(dataset: Dataset, ref_channel: Union[str, Mapping[int, str]], sink, tile_size=1000, overlap=100,
                                 num_pyr_lvl=4, num_iter=3, use_full_res_img=False, use_dog=False):
    """The per-cycle loop of __main__.py:320-437: cycle 1 passes through; every later cycle is registered to
    the registered previous one (serial dependency), and its flow warps all channels / z-planes."""
    ofreg = OptFlowRegistrator()
    ofreg.tile_size, ofreg.overlap = tile_size, overlap
    ofreg.num_pyr_lvl, ofreg.num_iterations = num_pyr_lvl, num_iter
    ofreg.use_full_res_img, ofreg.use_dog = use_full_res_img, use_dog
    warper = Warper()
    warper.tile_size, warper.overlap = tile_size, overlap

    sink = _rank0_sink(sink)
    cycles = list(dataset.keys())
    ncycles = len(cycles)
    ref_img: Optional[torch.Tensor] = None
    decisions = {}
    for cyc_id, cyc in enumerate(cycles):
        _say(f"Processing Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
        ref_ch = ref_channel[cyc] if isinstance(ref_channel, Mapping) else ref_channel
        zplanes = list(dataset[cyc][ref_ch].values())
        if cyc_id == 0:
            _say("Skipping as it is a reference image")
            ref_img = max_project_pages(zplanes)
            _say(f"Saving Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
            for ch, pages in dataset[cyc].items():
                for z_id, page in enumerate(pages.values()):
                    p = _load(page)
                    sink(cyc, ch, z_id, p.cpu().numpy() if isinstance(p, torch.Tensor) else np.asarray(p))
            continue
        mov_img = max_project_pages(zplanes)
        ofreg.ref_img, ofreg.mov_img = ref_img, mov_img      # device tensors: the flow stays on the GPU
        flow = ofreg.register()
        warper.image, warper.flow = mov_img, flow
        ref_img = warper.warp()                               # reference of the next cycle
        decisions[cyc] = ofreg.decisions
        _say(f"Saving Cycle {cyc} [{cyc_id + 1}/{ncycles}]")
        for ch, pages in dataset[cyc].items():
            warp_and_save_pages(sink, cyc, ch, flow, pages, tile_size, overlap)
        del flow
    return decisions


def run_opt_flow_reg(config: Union[str, Mapping], dataset: Dataset, sink, ref_channel=None):
    """run_opt_flow_reg (__main__.py:534-609) for an in-memory dataset: YAML -> parameters -> per-cycle loop."""
    if isinstance(config, str):
        import yaml
        with open(config) as f:
            config = yaml.safe_load(f)
    p = optflow_parameters(config)
    if ref_channel is None:
        ref_channel = config.get("Input", {}).get("ReferenceChannel")
    _say("Performing non-linear optical flow based image registration")
    register_and_save_ofreg_imgs(dataset, ref_channel, sink, p["tile_size"], p["overlap"], p["num_pyr_lvl"],
                                 p["num_iterations"], p["use_full_res_img"], p["use_dog"])
    _say("Finished\n")
