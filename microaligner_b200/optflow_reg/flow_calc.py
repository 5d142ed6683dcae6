"""TileFlowCalc: drop-in for the reference's optflow_reg/flow_calc.py:50-98.

calc_flow() runs the single-scale Farneback flow (prev = moving, next = reference, poly_n = 1,
poly_sigma = 1.7, Gaussian window) on every (tile_size + 2*overlap)^2 window and stitches the
centres -- or on the whole image when max(shape)/tile_size < 2, exactly the reference's switch."""
import numpy as np
import torch

from .. import ops


class TileFlowCalc:
    def __init__(self):
        self.ref_img = np.array([])
        self.mov_img = np.array([])
        self.num_iter = 1
        self.win_size = 51
        self.tile_size = 1000
        self.overlap = 100
        self.tile_range = None  # (first, last+1) row-major tile indices computed by this rank; None = all

    def calc_flow(self):
        host_result = not isinstance(self.ref_img, torch.Tensor)
        ref = ops.to_device(self.ref_img)
        mov = ops.to_device(self.mov_img, ref.device)
        untiled = max(ref.shape) / self.tile_size < 2
        self.ref_img = np.array([])
        self.mov_img = np.array([])
        flow = ops.farneback_tiles(mov, ref, 0 if untiled else self.tile_size, self.overlap, self.win_size,
                                   self.num_iter, None if untiled else self.tile_range)
        return ops.to_host(flow) if host_result else flow
