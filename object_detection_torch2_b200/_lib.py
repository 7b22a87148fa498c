"""ctypes binding of libssdhead.so (C ABI declared in include/ssdhead.h).

There is no CPU or eager fallback: if the shared library cannot be loaded the import fails loudly.
The library is built in-tree by ``object_detection_torch2_b200.build`` (nvcc, sm_100a).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libssdhead.so")


class ImageStats(ctypes.Structure):
    """Mirror of ``ssdh_image_stats`` (include/ssdhead.h)."""
    _fields_ = [("loss", c_float), ("thr_pos", c_float), ("thr_neg", c_float), ("pos_raw", c_int32),
                ("k_pos", c_int32), ("k_neg", c_int32), ("pos_sel", c_int32), ("neg_sel", c_int32)]


MAX_RANKS, XCHG_RING = 16, 256


class IpcHandle(ctypes.Structure):
    _fields_ = [("bytes", ctypes.c_ubyte * 64)]


class ScalarExchange(ctypes.Structure):
    """Mirror of ``ssdh_scalar_exchange`` (include/ssdhead.h)."""
    _fields_ = [("world", c_int32), ("rank", c_int32), ("ring", ctypes.c_uint32), ("reserved", ctypes.c_uint32),
                ("inbox", c_void_p * MAX_RANKS), ("counters", c_void_p)]


class LossOptions(ctypes.Structure):
    """Mirror of ``ssdh_loss_options`` (include/ssdhead.h)."""
    _fields_ = [("struct_bytes", ctypes.c_uint32), ("force_best_prior", c_int32), ("inputs_stable", c_int32), ("exact_math", c_int32),
                ("next_outputs", c_void_p), ("next_targets", c_void_p), ("ce_override", c_void_p), ("exchange", POINTER(ScalarExchange))]


# name -> (restype, argtypes); the single source the symbol test checks against the header
SIGNATURES = {
    "ssdh_version": (c_int, []),
    "ssdh_last_error": (c_char_p, []),
    "ssdh_device_info": (c_int, [POINTER(c_int)] * 4),
    "ssdh_default_boxes": (c_int, [c_void_p, c_void_p]),
    "ssdh_match": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p]),
    "ssdh_encode": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "ssdh_smooth_l1": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssdh_softmax_cross_entropy": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssdh_split_pos_neg": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "ssdh_kplus1_value": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ssdh_multibox_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ssdh_multibox_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssdh_multibox_loss_pipelined": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ssdh_multibox_loss_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, POINTER(LossOptions)]),
    "ssdh_scalar_exchange_bytes": (c_size_t, [c_int]),
    "ssdh_scalar_exchange_create": (c_int, [c_int, POINTER(c_void_p), POINTER(IpcHandle)]),
    "ssdh_scalar_exchange_open": (c_int, [POINTER(IpcHandle), POINTER(c_void_p)]),
    "ssdh_scalar_exchange_close": (c_int, [c_void_p]),
    "ssdh_scalar_exchange_destroy": (c_int, [c_void_p]),
    "ssdh_scalar_exchange_reduce": (c_int, [POINTER(ScalarExchange), c_int, c_void_p, c_void_p, c_void_p]),
    "ssdh_scale_inplace": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "ssdh_prefetch_l2": (c_int, [c_void_p, c_size_t, c_void_p]),
    "ssdh_expand_targets": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssdh_pack_head": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ssdh_unpack_head": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ssdh_pack_head_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "ssdh_unpack_head_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ssdh_decode": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "ssdh_score": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssdh_iou": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ssdh_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ssdh_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p,
                         c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssdh_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssdh_gather_detections": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "ssdh_eval_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ssdh_eval_accumulate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "ssdh_eval_accumulate_kept": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "ssdh_eval_status": (c_int, [c_void_p, POINTER(c_int), c_void_p]),
    "ssdh_voc_ap_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ssdh_voc_ap": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class SsdHeadError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load (building first if the .so is absent and nvcc exists).  Raises if neither is possible."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        try:
            _build.build()
        except Exception as exc:  # noqa: BLE001
            raise ImportError(
                f"libssdhead.so is missing at {LIB_PATH} and could not be built ({exc}). "
                "Run `python -m object_detection_torch2_b200.build`; there is no CPU fallback.") from exc
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().ssdh_last_error()
        raise SsdHeadError(f"{what} failed with code {code}: {msg.decode() if msg else ''}")
