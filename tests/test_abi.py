"""CPU: the C-ABI library loads and exports every symbol include/microaligner_b200.h declares;
argument validation that needs no GPU returns the documented error codes."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "microaligner_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ma_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from microaligner_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert _lib.lib.ma_version() >= 100


def test_argument_validation_without_gpu():
    from microaligner_b200 import _lib
    lib = _lib.lib
    assert lib.ma_pyrdown(None, 0, 10, 10, 0, None, 0, None) == -1
    assert b"ma_pyrdown" in lib.ma_last_error()
    assert lib.ma_warp_tiles(None, 0, 0, None, 1, 1, 10, 1, None, 0, None) == -1
    assert lib.ma_farneback_tiles(None, None, 0, 0, 1, 1, 1, 1, 1, 1, 0, 1, None, None, 0, None) == -1
    assert lib.ma_dog_u8(None, 0, 0, 30, 30, None, 0, None, None) == -1
    # workspace sizing is pure host arithmetic
    S, Sp = 1200, 1216
    assert lib.ma_farneback_workspace_bytes(5000, 5000, 1000, 100, 3) == 3 * 22 * S * Sp * 4
    assert lib.ma_farneback_workspace_bytes(640, 512, 0, 0, 1) == 22 * 640 * 512 * 4
    assert lib.ma_merge_workspace_bytes(2500, 3100, 1000) == 3 * 4 * 2 * 4
    assert lib.ma_nmi_workspace_bytes(10 ** 6, 10 ** 6) == 0
    assert lib.ma_nmi_workspace_bytes(4 * 10 ** 8, 10 ** 6) == 0      # histograms live in distributed shared memory


def test_host_api_surface():
    import numpy as np
    from microaligner_b200 import OptFlowRegistrator, Warper
    r = OptFlowRegistrator()
    assert (r.num_pyr_lvl, r.num_iterations, r.tile_size, r.overlap, r.use_full_res_img, r.use_dog) == (4, 3, 1000, 100, False, False)
    w = Warper()
    assert (w.tile_size, w.overlap) == (1000, 100) and len(w.image) == 0 and len(w.flow) == 0
    a, b = np.zeros((4, 4), np.uint8), np.ones((4, 4), np.uint8)
    r.ref_img, r.mov_img = a, b
    assert r.mov_img is a  # the reference's getter quirk (optflow_registrator.py:72-74)
    assert r.get_dog_sigmas(1) == (5, 9) and r.get_dog_sigmas(32) == (1, 2)
    assert r.dog(a, False) is a


def test_options():
    """ma_set_option validates its index (the option slots are reserved for A/B measurements of kernel variants)."""
    from microaligner_b200 import _lib
    assert _lib.lib.ma_set_option(0, 0) == 0
    assert _lib.lib.ma_set_option(99, 1) != 0
