"""Warper: drop-in for the reference's optflow_reg/warper.py:29-76 on the B200.

`image` (2-D uint8/uint16) and `flow` ((H, W, 2) float32) may be numpy arrays or CUDA tensors.
`warp()` remaps the image tile by tile exactly like the reference (cv.remap INTER_LINEAR with
map = tile-local grid - flow, zero outside the tile window) but without materialising tiles:
one kernel reads the image and flow and writes the stitched result.  Like the reference it
blanks `image` and `flow` afterwards.  The result is a numpy array when the image was one,
otherwise a CUDA tensor (device-resident hand-off inside the pipeline)."""
import numpy as np
import torch

from .. import ops, parallel
from ..engine import Engine


class Warper:
    def __init__(self):
        self.image = np.array([])
        self.flow = np.array([])
        self.tile_size = 1000
        self.overlap = 100
        self._slicer_info = {}
        # several ranks, CUDA tensors: True = every rank returns the whole image (its band is warped locally, the other
        # bands arrive over NVLink); False = the result is valid on Engine.warp_band() -- this rank's rows -- only
        self.gather_image = True

    def warp(self):
        image, flow = self.image, self.flow
        self.image = np.array([])           # the reference blanks its inputs (warper.py:41,45)
        self.flow = np.array([])
        comm = parallel.get()
        if isinstance(image, torch.Tensor):
            flow = ops.to_device(flow, image.device)
            return Engine(self.tile_size, self.overlap, comm=comm).warp(ops.to_device(image), flow, gather=self.gather_image)
        image = np.asarray(image)
        if image.ndim != 2 or image.dtype not in (np.uint8, np.uint16):
            raise TypeError(f"unsupported image: dtype {image.dtype}, {image.ndim} dimensions; expected 2-D uint8 or uint16")
        if tuple(flow.shape) != image.shape + (2,) or str(flow.dtype).replace("torch.", "") != "float32":
            raise ValueError(f"flow must be float32 of shape {image.shape + (2,)}, got {flow.dtype} {tuple(flow.shape)}")
        eng = Engine(self.tile_size, self.overlap, comm=comm)
        if comm.world == 1:
            flow_d = ops.to_device(flow)
            if image.shape[0] >= 2 * self.tile_size:
                # large host image: tile rows stream up / through the kernel / down on three streams
                return ops.warp_tiles_host_streamed(image, flow_d, self.tile_size, self.overlap)
            return ops.to_host(eng.warp(ops.to_device(image, flow_d.device), flow_d))
        # several ranks: each uploads the rows its tile windows read, warps its band and fills its rows of the
        # node-shared result.  A flow that is not already on the devices is uploaded the same way, rows only.
        if isinstance(flow, torch.Tensor):
            flow_d = flow
        else:
            flow_d = ops.mirrored(flow)
            if flow_d is None:
                flow_d, ev = ops.upload_rows(np.asarray(flow), eng.flow_rows_needed(image.shape))
                ops.wait_upload(ev)
        return eng.warp_host(image, flow_d)
