"""ctypes binding of libhgs_raster.so (the C ABI declared in include/hgs_raster.h).

The product path has NO fallback: if the CUDA library is missing or fails to load,
``lib()`` raises and every operator in this package fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libhgs_raster.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hgs_raster.h")

_p, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t

# name -> (restype, argtypes) ; order = include/hgs_raster.h
SIGNATURES = {
    "hgs_abi_version": (_i, []),
    "hgs_debug_launch_count": (C.c_ulonglong, []),
    "hgs_status_string": (C.c_char_p, [_i]),
    "hgs_project3d_fwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 4 + [_i] + [_p] * 6 + [_p]),
    "hgs_project3d_bwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 3 + [_p, _p, _i, _p, _i, _p, _i, _p, _ll] + [_p] * 3 + [_i, _p]),
    "hgs_project2d_fwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 3 + [_i] + [_p] * 6 + [_p]),
    "hgs_project2d_bwd": (_i, [_p] * 5 + [_i] * 4 + [_f] * 2 + [_p, _p, _i, _p, _i, _p, _i, _p, _i, _p, _ll] + [_p] * 3 + [_i, _p]),
    "hgs_sh_fwd": (_i, [_i, _i] + [_p] * 6 + [_ll, _p, _i, _i, _i] + [_p] + [_p]),
    "hgs_sh_bwd": (_i, [_i, _i] + [_p] * 6 + [_ll, _p, _p, _i, _i, _i, _i] + [_p] * 3 + [_i, _p]),
    "hgs_isect_count": (_i, [_p, _p, _ll, _i, _i, _i, _p, _p]),
    "hgs_scan_temp_bytes": (_sz, [_ll]),
    "hgs_isect_bin_temp_bytes": (_sz, [_ll, _i, _i, _i]),
    "hgs_isect_bin_bucket_bytes": (_sz, [_ll]),
    "hgs_exclusive_scan_i32": (_i, [_p, _p, _p, _ll, _p, _sz, _p]),
    "hgs_isect_emit": (_i, [_p] * 4 + [_i] * 5 + [_p, _p, _p]),
    "hgs_isect_bin_prepare": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "hgs_project3d_fwd_bin": (_i, [_p] * 5 + [_i] * 4 + [_f] * 4 + [_i] + [_p] * 9 + [_sz] +
                              [_i, _p, _i, _p, _p, _i, _p, _p] + [_p]),
    "hgs_isect_bin_scan": (_i, [_i, _i, _i, _i, _i, _p, _p, _sz, _p]),
    "hgs_isect_bin_sorted": (_i, [_p, _i, _i, _ll, _ll, _ll, _i, _i, _i] + [_p] * 4 + [_sz, _p, _sz, _p]),
    "hgs_isect_offset_encode": (_i, [_p, _ll, _i, _i, _i, _p, _p]),
    "hgs_blend3d_fwd": (_i, [_p] * 6 + [_i] * 6 + [_p, _p, _ll] + [_p] * 3 + [_p]),
    "hgs_blend3d_bwd": (_i, [_p] * 6 + [_i] * 6 + [_p, _p, _ll] + [_p] * 10 + [_p]),
    "hgs_blend3d_pack_bytes": (_sz, [_ll]),
    "hgs_blend3d_pack": (_i, [_p] * 7 + [_ll, _p, _ll, _i, _p, _p]),
    "hgs_blend3d_fwd_packed": (_i, [_p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 3 + [_p]),
    "hgs_blend3d_bwd_packed": (_i, [_p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 6 + [_p]),
    "hgs_gauss_bwd_fused": (_i, [_p, _i, _p, _ll, _i] + [_p] * 5 + [_i, _i, _f, _f, _f, _i, _i] + [_p] * 9 + [_p]),
    "hgs_blend3d_unpack": (_i, [_p, _p, _ll, _ll, _p, _p, _i, _p]),
    "hgs_zero_rows": (_i, [_p, _i, _p, _ll, _p]),
    "hgs_blend3d_stats": (_i, [_p, _i, _i, _i, _i, _p, _p, _ll, _p, _p]),
    "hgs_blend2d_pack_bytes": (_sz, [_ll]),
    "hgs_blend2d_pack": (_i, [_p] * 8 + [_ll, _ll, _i, _p, _p]),
    "hgs_blend2d_fwd_packed": (_i, [_p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 7 + [_p]),
    "hgs_blend2d_bwd_packed": (_i, [_p, _p] + [_i] * 6 + [_p, _p, _ll] + [_p] * 10 + [_p]),
    "hgs_normals_post_fwd": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _p, _p, _p]),
    "hgs_normals_post_bwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _p, _p, _i, _p]),
    "hgs_densify_stats": (_i, [_p, _i, _p, _p, _ll, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "hgs_l1_loss_partials": (_i, []),
    "hgs_l1_loss_fwd": (_i, [_p, _p, _p, _ll, _i, _f, _f, _p, _p, _p]),
    "hgs_l1_loss_bwd": (_i, [_p, _p, _p, _ll, _i, _f, _f, _p, _p, _p]),
    "hgs_ssim_partials": (_ll, [_i, _i, _i]),
    "hgs_ssim_fwd": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "hgs_ssim_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p]),
    "hgs_decode_count": (_i, [_p] * 5 + [_ll, _i, _i, _i, _i] + [_p] * 3 + [_p]),
    "hgs_decode_fwd": (_i, [_p] * 7 + [_ll, _i, _i, _i, _i, _i] + [_p] * 8 + [_p]),
    "hgs_decode_bwd": (_i, [_p] * 8 + [_ll, _i, _i, _i, _i, _i] + [_p] * 12 + [_p]),
    "hgs_anchor_filter": (_i, [_p, _p, _p, _p, _i, _p, _p, _f, _f, _f, _i, _i, _p, _p, _i, _i, _i, _f, _f, _f, _f, _p, _p]),
    "hgs_exchange_row_floats": (_i, [_p, _i]),
    "hgs_exchange_mailbox_bytes": (_sz, [_i, _ll, _ll, _i]),
    "hgs_exchange_push": (_i, [_p, _p, _i, _ll, _p, _ll, _ll, _p, _i, _i, C.c_ulonglong, _p]),
    "hgs_exchange_reduce": (_i, [_p, _p, _i, _ll, _ll, _p, _i, _i, C.c_ulonglong, _p, _p]),
    "hgs_exchange_vjp_mailbox_bytes": (_sz, [_i, _ll, _ll]),
    "hgs_exchange_vjp_push": (_i, [_i, _i] + [_p] * 9 + [_i, _i, _f, _f, _f, _ll, _p, _ll, _ll, _p, _p, _i, _i,
                                                     C.c_ulonglong, _p]),
    "hgs_exchange_vjp_push_2dgs": (_i, [_i, _i, _p, _i] + [_p] * 8 + [_i, _i, _f, _f, _ll, _p, _ll, _ll, _p, _p, _i, _i,
                                                                  C.c_ulonglong, _p]),
    "hgs_exchange_vjp_reduce": (_i, [_i, _i, _p, _ll, _ll, _p, _i, _i, C.c_ulonglong] + [_p] * 8 + [_p]),
    "hgs_peer_alloc": (_i, [_sz, C.POINTER(C.c_void_p)]),
    "hgs_peer_free": (_i, [_p]),
    "hgs_peer_export": (_i, [_p, _p]),
    "hgs_peer_import": (_i, [_p, C.POINTER(C.c_void_p)]),
    "hgs_peer_close": (_i, [_p]),
    "hgs_blend2d_fwd": (_i, [_p] * 7 + [_i] * 6 + [_p, _p, _ll] + [_p] * 7 + [_p]),
    "hgs_blend2d_bwd": (_i, [_p] * 7 + [_i] * 6 + [_p, _p, _ll] + [_p] * 16 + [_p]),
}

_lock = threading.Lock()
_lib = None


class HgsError(RuntimeError):
    pass


def declared_symbols() -> list:
    """Function names declared in include/hgs_raster.h (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hgs_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load (once) and return the C-ABI library; raise if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise HgsError(
                    f"{LIB_PATH} is missing: build it with `python -m horizongs_b200.csrc.build` "
                    "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for this path.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            if handle.hgs_abi_version() != 2:
                raise HgsError("libhgs_raster.so ABI version mismatch; rebuild")
            _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().hgs_status_string(status).decode()
        raise HgsError(f"{what} failed: {msg} (status {status})")


def ptr(t):
    """device pointer of a tensor as c_void_p (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())
