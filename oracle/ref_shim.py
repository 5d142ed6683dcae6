"""ORACLE (test infrastructure only -- never imported by the product path).

Import shim that lets the *unmodified* reference package (/root/reference/microaligner) run
in the build container, where dask / tifffile / scikit-image / pint are not installed.
Stub modules are injected into ``sys.modules`` before the import; the hot path only uses
``dask.delayed`` / ``dask.compute`` / ``dask.config.set``; the pipeline (``microaligner.__main__``) additionally uses a
small slice of ``tifffile`` -- served by the repo's own TIFF reader / writer -- and ``pint`` for one unit conversion.

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
    sys.modules["tifffile"] = _tifffile_stub()
    sys.modules["pint"] = _pint_stub()
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


def _tifffile_stub():
    """The slice of tifffile the reference's pipeline touches (TiffFile(...).series[0].{shape, axes, dtype, pages},
    .ome_metadata; memmap(path, shape=, dtype=, description=, bigtiff=True, contiguous=True)), served by the repo's own
    reader / writer (microaligner_b200/tiffio.py, itself cross-checked against libtiff in tests/test_tiffio_cpu.py)."""
    try:
        from microaligner_b200 import tiffio
    except ImportError:       # the CUDA library is not built: the optical-flow classes are still importable
        return types.ModuleType("tifffile")
    mod = types.ModuleType("tifffile")
    mod.TiffFile = tiffio.TiffFile

    def memmap(path, shape=None, dtype=None, description=None, **_ignored):
        return tiffio.memmap(path, shape, dtype, description=description)

    mod.memmap = memmap
    mod.imread = lambda path, key=0: tiffio.TiffFile(path).pages[key].asarray()
    return mod


def _pint_stub():
    """pint.UnitRegistry()[unit] * value -> .to("nm").magnitude, with exact SI factors (the only use: converting the
    physical pixel size of the OME-XML to nanometres, ome_meta_processing.py:47-54)."""
    nm = {"nm": 1.0, "um": 1e3, "µm": 1e3, "μm": 1e3, "micron": 1e3, "mm": 1e6, "cm": 1e7, "m": 1e9, "pm": 1e-3}

    class Q:
        def __init__(self, v, unit):
            self.magnitude, self.unit = v, unit

        def __rmul__(self, v):
            return Q(v * self.magnitude, self.unit)

        def to(self, unit):
            return Q(self.magnitude * nm[self.unit] / nm[unit], unit)

    class Reg:
        def __getitem__(self, unit):
            return Q(1.0, unit)

    mod = types.ModuleType("pint")
    mod.UnitRegistry = lambda *a, **k: Reg()
    return mod


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


def load_pipeline(workers: int = 1):
    """The reference's ``microaligner.__main__`` module (run_opt_flow_reg, register_and_save_ofreg_imgs, ...), its TIFF I/O
    served by the tifffile stub above."""
    load(workers)
    import importlib
    return importlib.import_module("microaligner.__main__")
