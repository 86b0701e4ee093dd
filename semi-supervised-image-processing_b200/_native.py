"""ctypes binding of libfx_b200.so (C ABI: include/fx_b200.h).

There is no fallback: if the CUDA library has not been built, or a call fails, this module raises.
Build with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C <pkg>/csrc``.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint8, c_uint32, c_uint64, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libfx_b200.so"

FX_OK = 0
FX_ERR_INVALID, FX_ERR_CUDA, FX_ERR_UNSUPPORTED, FX_ERR_NOMEM, FX_ERR_STATE = -1, -2, -3, -4, -5
PRECISION_BF16, PRECISION_FP32 = 0, 1
NUM_CONV_LAYERS = 20
EMBED_DIM = 512
CROP = 224
RESIZE = 256
MAX_LANES = 2
MAX_CLASSES = 32
TRANSFORM_EXTRACT, TRANSFORM_SQUARE224 = 0, 1
HOST_SLOTS = 4
JPEG_BACKEND_AUTO, JPEG_BACKEND_HARDWARE, JPEG_BACKEND_GPU = 0, 1, 2
FILE_GPU_JPEG, FILE_HOST_DECODE, FILE_UNREADABLE = 0, 1, 2

# every symbol include/fx_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = (
    "fx_version", "fx_abi_version", "fx_last_error", "fx_create", "fx_destroy", "fx_load_weights",
    "fx_preprocess_nchw_f32", "fx_preprocess", "fx_forward", "fx_stage_nchw_f32", "fx_select_lane", "fx_set_transform", "fx_load_head", "fx_classify", "fx_embed", "fx_embed_host", "fx_embed_host_async", "fx_embed_host_async_dev", "fx_embed_host_wait",
    "fx_jpeg_init", "fx_jpeg_backend", "fx_jpeg_probe", "fx_jpeg_read_files", "fx_jpeg_decode", "fx_embed_files_async",
    "fx_column_stats", "fx_standardize", "fx_neighbor_probe", "fx_launch_count", "fx_h2d_bytes", "fx_profile_enable", "fx_profile_read", "fx_debug_conv", "fx_debug_folded", "fx_debug_tma_probe", "fx_debug_umma_shift", "fx_debug_stem_pool", "fx_debug_mma_rate", "fx_debug_staging",
    "fx_host_resized_size", "fx_host_crop_offset", "fx_host_coeffs",
)


class FxError(RuntimeError):
    """A failed call into libfx_b200.so (status code + the library's error text)."""

    def __init__(self, status: int, text: str):
        super().__init__(f"libfx_b200 status {status}: {text}")
        self.status = status
        self.text = text


class ImageDesc(ctypes.Structure):
    _fields_ = [("offset", c_uint64), ("height", c_int32), ("width", c_int32), ("channels", c_int32), ("reserved", c_int32)]


class FileInfo(ctypes.Structure):
    _fields_ = [("status", c_int32), ("height", c_int32), ("width", c_int32), ("components", c_int32), ("subsampling", c_int32),
                ("encoding", c_int32), ("precision", c_int32), ("reserved", c_int32), ("offset", c_uint64), ("length", c_uint64)]


class MatrixStats(ctypes.Structure):
    _fields_ = [("nan_count", ctypes.c_int64), ("inf_count", ctypes.c_int64), ("mean_abs_mean", ctypes.c_double), ("mean_std", ctypes.c_double)]


class ConvBn(ctypes.Structure):
    _fields_ = [
        ("weight", POINTER(c_float)), ("gamma", POINTER(c_float)), ("beta", POINTER(c_float)),
        ("mean", POINTER(c_float)), ("var", POINTER(c_float)), ("eps", c_float),
        ("cout", c_int32), ("cin", c_int32), ("kh", c_int32), ("kw", c_int32), ("stride", c_int32), ("pad", c_int32),
    ]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repository root. "
            "There is no CPU or PyTorch fallback for this path."
        )
    L = ctypes.CDLL(str(LIB_PATH))
    L.fx_version.restype = c_char_p
    L.fx_abi_version.restype = c_int
    L.fx_last_error.restype = c_char_p
    L.fx_last_error.argtypes = [c_void_p]
    L.fx_create.argtypes = [POINTER(c_void_p), c_int, c_int, c_int]
    L.fx_destroy.argtypes = [c_void_p]
    L.fx_destroy.restype = None
    L.fx_load_weights.argtypes = [c_void_p, POINTER(ConvBn), c_int]
    L.fx_preprocess_nchw_f32.argtypes = [c_void_p, c_void_p, POINTER(ImageDesc), c_int, c_void_p, c_void_p]
    L.fx_preprocess.argtypes = [c_void_p, c_void_p, POINTER(ImageDesc), c_int, c_void_p]
    L.fx_forward.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.fx_stage_nchw_f32.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.fx_select_lane.argtypes = [c_void_p, c_int]
    L.fx_set_transform.argtypes = [c_void_p, c_int]
    L.fx_load_head.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
    L.fx_classify.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.fx_embed.argtypes = [c_void_p, c_void_p, POINTER(ImageDesc), c_int, c_void_p, c_void_p]
    L.fx_embed_host.argtypes = [c_void_p, c_void_p, c_size_t, POINTER(ImageDesc), c_int, c_void_p]
    L.fx_embed_host_async.argtypes = [c_void_p, c_int, c_void_p, c_size_t, POINTER(ImageDesc), c_int, c_void_p]
    L.fx_embed_host_async_dev.argtypes = [c_void_p, c_int, c_void_p, c_size_t, POINTER(ImageDesc), c_int, c_void_p]
    L.fx_embed_host_wait.argtypes = [c_void_p, c_int]
    L.fx_jpeg_init.argtypes = [c_void_p, c_int]
    L.fx_jpeg_backend.argtypes = [c_void_p]
    L.fx_jpeg_probe.argtypes = [c_void_p, c_void_p, c_size_t, POINTER(FileInfo)]
    L.fx_jpeg_read_files.argtypes = [c_void_p, c_int, POINTER(c_char_p), c_int, POINTER(FileInfo)]
    L.fx_jpeg_decode.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_size_t), c_int, c_void_p, POINTER(ImageDesc), c_void_p]
    L.fx_embed_files_async.argtypes = [c_void_p, c_int, POINTER(FileInfo), POINTER(c_void_p), POINTER(ImageDesc), c_int, c_size_t, c_void_p, c_void_p]
    L.fx_column_stats.argtypes = [c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p, c_void_p, c_void_p, POINTER(MatrixStats), c_void_p]
    L.fx_standardize.argtypes = [c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.fx_neighbor_probe.argtypes = [c_void_p, c_void_p, ctypes.c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    L.fx_launch_count.argtypes = [c_void_p]
    L.fx_launch_count.restype = c_uint64
    L.fx_h2d_bytes.argtypes = [c_void_p]
    L.fx_h2d_bytes.restype = c_uint64
    L.fx_profile_enable.argtypes = [c_void_p, c_int]
    L.fx_profile_read.argtypes = [c_void_p, c_void_p, c_int]
    L.fx_debug_conv.argtypes = [c_void_p, POINTER(ConvBn), c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.fx_debug_folded.argtypes = [c_void_p, c_int, c_void_p, c_void_p]
    L.fx_debug_tma_probe.argtypes = [c_void_p, c_void_p, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint32),
                                     POINTER(c_uint32), c_int, POINTER(c_int), c_int, c_void_p]
    L.fx_debug_stem_pool.argtypes = [c_void_p, POINTER(ConvBn), c_void_p, c_int, c_void_p, c_void_p]
    L.fx_debug_staging.argtypes = [c_void_p, c_int, c_void_p, c_size_t, c_void_p]
    L.fx_debug_mma_rate.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.fx_debug_umma_shift.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.fx_host_resized_size.argtypes = [c_int, c_int, POINTER(c_int), POINTER(c_int)]
    L.fx_host_crop_offset.argtypes = [c_int]
    L.fx_host_coeffs.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_int]
    if L.fx_abi_version() != 1:
        raise RuntimeError(f"{LIB_PATH} has ABI version {L.fx_abi_version()}, this binding expects 1; rebuild it")
    _lib = L
    return L


def check(status: int, handle=None) -> None:
    if status != FX_OK:
        text = lib().fx_last_error(handle)
        raise FxError(status, text.decode("utf-8", "replace") if text else "")
