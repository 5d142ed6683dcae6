"""Seeded synthetic inputs shared by the tests, the golden generator and bench.py."""
import numpy as np


from benchdata import synth_pair  # noqa: F401  (re-exported)


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
