"""Seeded synthetic image pairs for bench.py and the tests: a smooth random texture
`ref` and `mov` = ref displaced by a known sinusoidal field (amplitude 3 px, period 512 px).

Host-side data generator only (numpy + cv2); nothing here is on the product path.  Large images are
produced block-wise (cv2.remap addresses its source with int16 coordinates, and 50 000^2 float maps
would not fit comfortably in host RAM)."""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np


def synth_pair(h, w, seed=0, dtype=np.uint16, amp=3.0, period=512.0):
    import cv2
    rng = np.random.default_rng(seed)
    n = rng.random((h, w), dtype=np.float32)
    b = cv2.GaussianBlur(n, (0, 0), 3)
    b = (b - b.min()) / (b.max() - b.min())
    full = 65535 if dtype == np.uint16 else 255
    ref = (b * 0.9 * full).astype(dtype)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    dx = amp * np.sin(2 * np.pi * y / period)
    dy = 0.66 * amp * np.cos(2 * np.pi * x / period)
    mov = cv2.remap(ref, (x + dx).astype(np.float32), (y + dy).astype(np.float32), cv2.INTER_LINEAR)
    return ref, mov


def synth_pair_large(h, w, seed=0, dtype=np.uint16, amp=3.0, period=512.0, block=2048, workers=None, out=None):
    """Same distribution as synth_pair, generated in blocks with a thread pool.
    `out` = optional (ref, mov) pre-allocated (e.g. page-locked) arrays to fill."""
    import cv2
    workers = workers or min(32, os.cpu_count() or 1)
    full = 65535 if dtype == np.uint16 else 255
    ref = out[0] if out is not None else np.empty((h, w), dtype)
    mov = out[1] if out is not None else np.empty((h, w), dtype)
    halo = 16
    blocks = [(y0, x0) for y0 in range(0, h, block) for x0 in range(0, w, block)]

    def noise_block(y0, x0):
        # noise is a pure function of (seed, block index): halos are regenerated consistently
        return np.random.default_rng([seed, y0 // block, x0 // block]).random(
            (min(block, h - y0), min(block, w - x0)), dtype=np.float32)

    def make_ref(yx):
        y0, x0 = yx
        bh, bw = min(block, h - y0), min(block, w - x0)
        ya, yb = max(y0 - halo, 0), min(y0 + bh + halo, h)
        xa, xb = max(x0 - halo, 0), min(x0 + bw + halo, w)
        tile = np.empty((yb - ya, xb - xa), np.float32)
        for by in range(ya // block * block, yb, block):
            for bx in range(xa // block * block, xb, block):
                nb = noise_block(by, bx)
                sy0, sy1 = max(ya, by), min(yb, by + nb.shape[0])
                sx0, sx1 = max(xa, bx), min(xb, bx + nb.shape[1])
                tile[sy0 - ya:sy1 - ya, sx0 - xa:sx1 - xa] = nb[sy0 - by:sy1 - by, sx0 - bx:sx1 - bx]
        bl = cv2.GaussianBlur(tile, (0, 0), 3)[y0 - ya:y0 - ya + bh, x0 - xa:x0 - xa + bw]
        # sigma-3 blur of U(0,1) noise has a narrow, known range: fixed affine map instead of a global min/max
        v = np.clip((bl - 0.35) / 0.30, 0.0, 1.0)
        ref[y0:y0 + bh, x0:x0 + bw] = (v * 0.9 * full).astype(dtype)

    def make_mov(yx):
        y0, x0 = yx
        bh, bw = min(block, h - y0), min(block, w - x0)
        pad = int(np.ceil(amp)) + 2
        ya, yb = max(y0 - pad, 0), min(y0 + bh + pad, h)
        xa, xb = max(x0 - pad, 0), min(x0 + bw + pad, w)
        yy, xx = np.mgrid[y0:y0 + bh, x0:x0 + bw].astype(np.float32)
        mx = (xx + amp * np.sin(2 * np.pi * yy / period) - xa).astype(np.float32)
        my = (yy + 0.66 * amp * np.cos(2 * np.pi * xx / period) - ya).astype(np.float32)
        mov[y0:y0 + bh, x0:x0 + bw] = cv2.remap(np.ascontiguousarray(ref[ya:yb, xa:xb]), mx, my, cv2.INTER_LINEAR,
                                                borderMode=cv2.BORDER_REPLICATE)

    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(make_ref, blocks))
        list(ex.map(make_mov, blocks))
    return ref, mov
