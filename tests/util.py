"""Seeded synthetic inputs shared by the tests, the golden generator and bench.py."""
import numpy as np


def synth_pair(h, w, seed=0, dtype=np.uint16, amp=3.0, period=512.0):
    """Smooth random texture `ref` and `mov` = ref displaced by a sinusoidal field of amplitude `amp`."""
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


def blobs(h, w, seed=0, dtype=np.uint16):
    """Microscopy-like: sparse Gaussian blobs on a dark background + Poisson noise."""
    import cv2
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float32)
    n = max(1, (h * w) // 1600)
    ys, xs = rng.integers(0, h, n), rng.integers(0, w, n)
    img[ys, xs] = rng.uniform(2000, 40000, n).astype(np.float32)
    img = cv2.GaussianBlur(img, (0, 0), 6) * 60 + 100
    img = rng.poisson(np.clip(img, 0, None)).astype(np.float32)
    full = 65535 if dtype == np.uint16 else 255
    if dtype == np.uint8:
        img = img / max(img.max(), 1) * 255
    return np.clip(img, 0, full).astype(dtype)


def random_flow(h, w, seed=0, mag=4.0):
    import cv2
    rng = np.random.default_rng(seed)
    f = rng.standard_normal((h, w, 2)).astype(np.float32)
    f = cv2.GaussianBlur(f, (0, 0), 15) * 15 * mag
    return np.ascontiguousarray(f.astype(np.float32))
