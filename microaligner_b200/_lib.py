"""ctypes binding of libmicroaligner_b200.so (C ABI declared in include/microaligner_b200.h).

There is no CPU fallback: if the shared library is missing this module raises at import and every
operator of the package is unusable.  Build it with ``python build.py`` (nvcc, sm_100a)."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmicroaligner_b200.so")

MA_U8, MA_U16, MA_F32 = 0, 1, 2


class MicroalignerB200Error(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
        "Run `python build.py` at the repository root to compile it for sm_100a.")

lib = ctypes.CDLL(LIB_PATH)

# every symbol include/microaligner_b200.h declares, with its prototype
PROTOTYPES = {
    "ma_version": (c_int, []),
    "ma_last_error": (c_char_p, []),
    "ma_set_option": (c_int, [c_int, c_int]),
    "ma_pyrdown": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ma_pyrdown_rows": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "ma_pyrup_flow": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_void_p]),
    "ma_warp_tiles": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ma_warp_tiles_rows": (c_int, [c_void_p, c_size_t, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "ma_warp_affine": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_double), c_void_p, c_size_t,
                               c_int, c_int, c_void_p]),
    "ma_pyrup_flow_rows": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_float, c_int, c_int, c_void_p]),
    "ma_merge_flows_tile_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ma_dog_diff_pitch_floats": (c_size_t, [c_int]),
    "ma_dog_diff_rows": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ma_dog_band_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ma_dog_quantize_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ma_nmi_chunk_range": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ma_nmi_chunk_range2": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ma_compose_flows_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "ma_merge_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ma_merge_flows_tiles": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ma_farneback_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "ma_farneback_tiles": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ma_farneback_tiles_ex": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p, c_void_p, c_size_t, ctypes.c_uint, c_void_p]),
    "ma_dog_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ma_dog_u8": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "ma_nmi_workspace_bytes": (c_size_t, [c_size_t, c_size_t]),
    "ma_nmi_chunks": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ma_zmip_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ma_zmip_normalize_u8": (c_int, [ctypes.POINTER(c_void_p), c_int, c_size_t, c_int, c_int, c_int, c_void_p, c_size_t,
                                     c_void_p, c_void_p]),
    "ma_minmax": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ma_tiff_lzw_decode": (ctypes.c_longlong, [c_char_p, c_size_t, c_void_p, c_size_t]),
    "ma_launch_count": (ctypes.c_longlong, []),
    "ma_profile_kernels": (c_int, []),
    "ma_profile_kernel_name": (c_char_p, [c_int]),
    "ma_profile_enable": (None, [c_int]),
    "ma_profile_reset": (None, []),
    "ma_profile_read": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_double)]),
}


def profile_summary():
    """{kernel name: (total ms, launches, units)} accumulated since ma_profile_reset()."""
    out = {}
    for i in range(lib.ma_profile_kernels()):
        ms, n, u = c_double(), ctypes.c_longlong(), c_double()
        lib.ma_profile_read(i, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(u))
        if n.value:
            out[lib.ma_profile_kernel_name(i).decode()] = (ms.value, n.value, u.value)
    return out

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args



def check(status: int, what: str):
    if status != 0:
        msg = lib.ma_last_error()
        raise MicroalignerB200Error(f"{what} failed ({status}): {msg.decode() if msg else ''}")
