"""Similarity gate of the registration loop (reference shared_modules/similarity_scoring.py:27-68).

mi_tiled: NMI of the whole image when max(shape)/tile_size < 2, otherwise the unweighted mean of
the NMI of consecutive tile_size^2-element chunks of the row-major flattened images.  The joint
histograms and entropies are computed on the device (ma_nmi_chunks); the mean over the handful of
per-chunk doubles is taken with numpy on the host so it rounds exactly like the reference's
np.mean."""
from typing import List, Tuple

import numpy as np
import torch

from .. import ops


def mi_tiled(arr1: torch.Tensor, arr2: torch.Tensor, tile_size: int) -> float:
    if max(arr1.shape) / tile_size < 2:
        scores = ops.nmi_chunks(arr1, arr2, arr1.numel())
        return float(scores.cpu().numpy()[0])
    scores = ops.nmi_chunks(arr1, arr2, tile_size * tile_size)
    return float(np.mean(scores.cpu().numpy()))


def mutual_information_test(ref_arr, test_arr, init_arr, tile_size: int) -> Tuple[float, float]:
    after_mi_score = mi_tiled(ref_arr, test_arr, tile_size)
    before_mi_score = mi_tiled(ref_arr, init_arr, tile_size)
    return after_mi_score, before_mi_score


def check_if_higher_similarity(ref_arr, test_arr, init_arr, tile_size: int) -> List[bool]:
    mi_scores = mutual_information_test(ref_arr, test_arr, init_arr, tile_size)
    print("    MI score after:", mi_scores[0], "| MI score before:", mi_scores[1])
    return [mi_scores[0] > mi_scores[1]]
