"""ORACLE (test infrastructure only -- never imported by the product path).

Import shim that lets the *unmodified* reference package (/root/reference/microaligner) run
in the build container, where dask / tifffile / scikit-image / pint are not installed.
Stub modules are injected into ``sys.modules`` before the import; the hot path only uses
``dask.delayed`` / ``dask.compute`` / ``dask.config.set``.

Used only to (a) pin oracle/reference_flow.py against the real reference and (b) generate
tests/golden/*.npz (scripts/make_golden.py).  /root/reference does not exist on the GPU box:
nothing in `-m gpu` tests, smoke() or bench.py calls this.
"""
import os
import sys
import types
from concurrent.futures import ThreadPoolExecutor

REFERENCE_ROOT = os.environ.get("MICROALIGNER_REFERENCE", "/root/reference")


class _Delayed:
    def __init__(self, fn, args, kwargs):
        self.fn, self.args, self.kwargs = fn, args, kwargs

    def run(self):
        return self.fn(*self.args, **self.kwargs)


def _install_stubs(workers: int):
    dask = types.ModuleType("dask")
    dask.delayed = lambda fn: (lambda *a, **k: _Delayed(fn, a, k))

    def compute(*tasks):
        if workers <= 1:
            return tuple(t.run() for t in tasks)
        with ThreadPoolExecutor(workers) as ex:
            return tuple(ex.map(lambda t: t.run(), tasks))

    dask.compute = compute
    dask.config = types.SimpleNamespace(set=lambda *a, **k: None)
    sys.modules["dask"] = dask
    for name in ("tifffile", "pint"):
        sys.modules.setdefault(name, types.ModuleType(name))
    pint = sys.modules["pint"]
    if not hasattr(pint, "UnitRegistry"):
        pint.UnitRegistry = lambda *a, **k: None
    sk = types.ModuleType("skimage")
    skt = types.ModuleType("skimage.transform")
    skt.AffineTransform = None
    skt.warp = None
    sk.transform = skt
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.transform", skt)
    cv2 = __import__("cv2")
    if not hasattr(cv2, "xfeatures2d"):  # feature_reg imports cv2 only; attribute used lazily
        pass


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "microaligner"))


def load(workers: int = 1):
    """Return the reference's ``microaligner.optflow_reg`` module (OptFlowRegistrator, Warper)."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _install_stubs(workers)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    return importlib.import_module("microaligner.optflow_reg")
