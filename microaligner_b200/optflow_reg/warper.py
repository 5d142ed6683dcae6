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

    def warp(self):
        host_result = not isinstance(self.image, torch.Tensor)
        if host_result and parallel.get().world == 1 and isinstance(self.image, np.ndarray) and self.image.ndim == 2 \
                and self.image.dtype in (np.uint8, np.uint16) and self.image.shape[0] >= 2 * self.tile_size:
            # large host image, single GPU: stream tile rows up / through the kernel / down on three streams
            flow = ops.to_device(self.flow)
            image, self.image, self.flow = self.image, np.array([]), np.array([])
            if tuple(flow.shape) != image.shape + (2,) or flow.dtype != torch.float32:
                raise ValueError(f"flow must be float32 of shape {image.shape + (2,)}, got {flow.dtype} {tuple(flow.shape)}")
            return ops.warp_tiles_host_streamed(image, flow, self.tile_size, self.overlap)
        img = ops.to_device(self.image, self.flow.device if isinstance(self.flow, torch.Tensor) else None)
        flow = ops.to_device(self.flow, img.device)
        self.image = np.array([])
        self.flow = np.array([])
        out = Engine(self.tile_size, self.overlap, comm=parallel.get()).warp(img, flow)
        return ops.to_host(out) if host_result else out

    def warp_sharded(self):
        """Opt-in companion of OptFlowRegistrator.register_sharded() for several ranks: `image` is a full-shape host array
        (only the rows this rank's tile windows read are uploaded), `flow` the device flow `reg.device_flow`.
        Returns (rows, warped): rows [rows[0], rows[1]) of the warped image, computed and downloaded by this rank."""
        if not isinstance(self.flow, torch.Tensor):
            raise TypeError("warp_sharded() needs the device flow of register_sharded() (reg.device_flow)")
        image, flow = np.asarray(self.image), self.flow
        self.image = np.array([])
        self.flow = np.array([])
        if tuple(flow.shape) != image.shape + (2,) or flow.dtype != torch.float32:
            raise ValueError(f"flow must be float32 of shape {image.shape + (2,)}, got {flow.dtype} {tuple(flow.shape)}")
        return Engine(self.tile_size, self.overlap, comm=parallel.get()).warp_host_sharded(image, flow)
