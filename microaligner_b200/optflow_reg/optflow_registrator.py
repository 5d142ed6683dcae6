"""OptFlowRegistrator: drop-in for the reference's optflow_reg/optflow_registrator.py:50-274.

Same attribute surface (ref_img, mov_img, num_pyr_lvl, num_iterations, tile_size, overlap,
use_full_res_img, use_dog), same ValueErrors, same stdout lines, same coarse-to-fine control flow
-- including the reference's quirks -- but every pixel stays in HBM from
the first upload to the returned flow: pyramids, DoG images, flows and the similarity histograms
are device tensors, and per pyramid level only the per-chunk NMI doubles cross to the host.

Inputs may be numpy arrays (result: numpy (H, W, 2) float32, like the reference) or CUDA tensors
(result: CUDA tensor, for a device-resident hand-off to Warper).

The loop itself -- pyramids, pre-warp, [DoG], tiled Farneback, post-warp, NMI gate, merge / up-sample, written
once for 1..N GPUs -- lives in microaligner_b200/engine.py; this class is the reference-shaped front door."""
from typing import List, Tuple

import numpy as np
import torch

from .. import ops, parallel
from ..engine import Engine
from ..shared_modules.img_checks import check_img_dims_match, check_img_is_2d_grey, check_img_is_provided


class OptFlowRegistrator:
    def __init__(self):
        self._ref_img = np.array([])
        self._mov_img = np.array([])
        self.num_pyr_lvl = 4
        self.num_iterations = 3
        self.tile_size = 1000
        self.overlap = 100
        self.use_full_res_img = False
        self.use_dog = False
        self.decisions: List[dict] = []  # per level: factor, mi_after, mi_before, better (diagnostic)
        # numpy results are READ-ONLY and stay mirrored on the device: handing the flow to Warper.flow then needs no
        # upload.  The mirror holds the device copy of the flow (8 bytes per pixel of HBM, on several ranks this rank's
        # band of it) for as long as the returned array is alive -- keep one flow per cycle alive and that many flows
        # stay in HBM.  Set False for a plain writeable array without a device copy, as the reference returns.
        self.mirror_flow = True
        # multi-GPU only (parallel.init): False leaves a device-resident flow sharded -- each rank's tensor is
        # valid on its own band of tile rows, which is all Warper.warp() on the same ranks needs
        self.gather_flow = True
        # True (default): the Farneback window blur rounds multiply and add separately like OpenCV's CPU code, the
        # flow is bit-identical to the reference's.  False: FMA-contracted blur, ~1.5x faster in the two dominant
        # kernels, each Farneback call within ~1e-6 px of the exact one (inside the 0.01 / 0.1 px contract)
        self.exact_arithmetic = True
        # False (default) reproduces the reference including its quirks.  True: flows of successive levels are
        # composed properly (m = f2 + f1(p - f2) in image coordinates) and the final up-sampling is scaled by 2 --
        # NOT what the reference computes, but much closer to the true displacement
        self.corrected_composition = False

    @property
    def ref_img(self):
        return self._ref_img

    @ref_img.setter
    def ref_img(self, img):
        check_img_is_2d_grey(img, "ref")
        self._ref_img = img

    @property
    def mov_img(self):
        # the reference's getter returns the *reference* image (optflow_registrator.py:72-74); kept
        return self._ref_img

    @mov_img.setter
    def mov_img(self, img):
        check_img_is_2d_grey(img, "mov")
        self._mov_img = img

    # ------------------------------------------------------------------ helpers
    def get_dog_sigmas(self, pyr_factor: int) -> Tuple[int, int]:
        if pyr_factor > 16:
            return 1, 2
        return {1: (5, 9), 2: (4, 7), 4: (3, 5), 8: (2, 3), 16: (1, 2)}[pyr_factor]

    def dog(self, img, use_it: bool, low_sigma: int = 5, high_sigma: int = 9):
        """Difference of Gaussians (sigma 5 / 9, 41-tap) -> uint8; identity when use_it is False.
        An all-zero image yields all-zero labels (the reference returns the zero input itself)."""
        if not use_it:
            return img
        if (low_sigma, high_sigma) != (5, 9):
            raise NotImplementedError("the device DoG implements the reference's fixed sigmas (5, 9)")
        if isinstance(img, torch.Tensor):
            return ops.dog_u8(img)
        return ops.to_host(ops.dog_u8(ops.to_device(img)))

    # ------------------------------------------------------------------ the hot path
    def _engine(self) -> Engine:
        return Engine(self.tile_size, self.overlap, self.num_pyr_lvl, self.num_iterations, self.use_full_res_img,
                      self.use_dog, comm=parallel.get(), contract_fma=not self.exact_arithmetic,
                      corrected=self.corrected_composition)

    def register(self):
        """(H, W, 2) float32 flow of mov_img onto ref_img (optflow_registrator.py:93-173).

        numpy images -> numpy flow.  Host traffic is kept off the critical path: the images go up on a copy stream, the
        rows of the flow come down while the device is still working on the rest.  Under a process group
        (parallel.init, one process per GPU) every rank passes full-shape images, of which it reads and uploads only its
        own rows, and every rank gets the full flow back: the array lives in node-shared memory and each rank fills in the
        rows it computed over its own PCIe link.
        CUDA tensors -> CUDA tensor (device-resident hand-off to Warper; see `gather_flow` for several ranks)."""
        check_img_is_provided(self._ref_img, "ref")
        check_img_is_provided(self._mov_img, "mov")
        check_img_dims_match(self._ref_img, self._mov_img)
        eng = self._engine()
        if isinstance(self._ref_img, torch.Tensor):
            ref = ops.to_device(self._ref_img)
            mov = ops.to_device(self._mov_img, ref.device)
            eng.gather_flow = self.gather_flow
            m_flow = eng.register(ref, mov)     # the coarse-to-fine loop, sharded over the ranks of parallel.get()
            self.decisions = eng.decisions
            return m_flow
        ref, mov = np.asarray(self._ref_img), np.asarray(self._mov_img)
        for a in (ref, mov):
            if a.dtype not in (np.uint8, np.uint16):
                raise TypeError(f"unsupported image dtype {a.dtype}; expected uint8 or uint16")
        flow, dev_flow = eng.register_host(ref, mov)
        self.decisions = eng.decisions
        if self.mirror_flow:
            ops.mirror(flow, dev_flow)
        return flow
