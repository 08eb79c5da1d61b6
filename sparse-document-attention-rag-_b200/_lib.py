"""ctypes binding of libsdag_b200.so (the C ABI declared in include/sdag_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import logging
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SDAG_B200_LIB: A/B measurements against another build of the same sources (e.g. an earlier round's); symbols that build
# lacks are then skipped at load time instead of failing
LIB_PATH = os.environ.get("SDAG_B200_LIB") or os.path.join(_HERE, "libsdag_b200.so")

SDAG_TILE = 128
BLOCKMAP_HEADER = 16
LAYOUT_HEADER = 8
MAX_NBR_K = 16
MAX_DOCS = 512


class SdagB200Error(RuntimeError):
    pass


_i32, _i64, _f32, _sz, _vp = C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_void_p

# name -> (restype, argtypes); every symbol include/sdag_b200.h declares
SIGNATURES = {
    "sdag_b200_version": (_i32, []),
    "sdag_b200_error_string": (C.c_char_p, [_i32]),
    "sdag_layout_ints": (_i32, [_i32, _i32]),
    "sdag_build_mask": (_i32, [_vp, _i32, _i32, _vp, _i32, _vp]),
    "sdag_blockmap_bytes": (_sz, [_i32, _i32, _i32]),
    "sdag_build_blockmap": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _sz, _vp]),
    "sdag_maskmap_bytes": (_sz, [_i32]),
    "sdag_blockmap_from_mask": (_i32, [_vp, _i64, _i32, _vp, _sz, _vp]),
    "sdag_prefill_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "sdag_prefill_fwd": (_i32, [_vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64,
                                _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _f32,
                                _vp, _sz, _i32, _vp, _sz, _i32, _vp]),
    "sdag_decode_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "sdag_decode_fwd": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _vp,
                               _i32, _i32, _i32, _i32, _i32, _f32, _vp, _sz, _vp, _i32, _vp]),
    "ip_cast_rows_bf16": (_i32, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "ip_search_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "ip_search": (_i32, [_vp, _i32, _vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _sz, _i32, _vp]),
    "topk_merge": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "topk_exchange_buffer_bytes": (_sz, [_i32, _i32, _i32]),
    "topk_exchange_merge": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, C.c_uint32, _vp, _vp, _vp]),
    "topk_pack": (_i32, [_vp, _vp, _i64, _vp, _vp]),
    "topk_merge_packed": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "centroid_select": (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "doc_knn": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "llm_rmsnorm": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "llm_rope_qkv": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _i32, _vp, _i32, _vp]),
    "llm_swiglu": (_i32, [_vp, _vp, _i64, _i32, _vp]),
    "llm_gemv": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "llm_skinny_linear": (_i32, [_vp, _i64, _vp, _vp, _vp, _f32, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "llm_prefetch_l2": (_i32, [_vp, _sz, _vp]),
    "llm_gemm_rope": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _i32, _vp]),
    "llm_gemv_rope": (_i32, [_vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _i32, _vp,
                              _i32, _vp]),
    "llm_gemm": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "llm_gemm_residual": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp]),
    "sdag_full_blockmap_bytes": (_sz, [_i32]),
    "sdag_build_blockmap_full": (_i32, [_vp, _i32, _i32, _vp, _sz, _vp]),
    "enc_layernorm": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp]),
    "enc_gelu": (_i32, [_vp, _vp, _i64, _vp]),
    "enc_mean_pool_normalize": (_i32, [_vp, _i64, _i64, _vp, _i32, _i32, _i32, _vp, _i64, _vp, _vp, _i32, _vp]),
}

_lib = None


def load():
    """Load the shared library once; raise if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdagB200Error(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if os.environ.get("SDAG_B200_LIB") and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().sdag_b200_error_string(code).decode()
        raise SdagB200Error(f"{what} failed: {msg} (code {code})")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> 0)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    """Raw ``cudaStream_t`` of torch's current stream on ``device``.  The library launches on the CUDA *current* device
    (tensor maps, ``cudaFuncSetAttribute``, the SM count are per device), so the tensor's device is made current here —
    every launch site fetches its stream through this function right before the call."""
    import torch

    if device is not None:
        dev = torch.device(device)
        if dev.type == "cuda" and dev.index is not None and dev.index != torch.cuda.current_device():
            torch.cuda.set_device(dev)
    return torch.cuda.current_stream(device).cuda_stream


# Every place the host code leaves its fast path (fused engine unavailable for a model, CUDA-graph capture failed, ...)
# reports here: counted, logged once per reason, and an error under SDAG_STRICT=1 — so a 3x performance cliff cannot
# pass silently (bench.py prints the table; tests/bench run strict).
fallbacks: dict = {}
_log = logging.getLogger("sdag_b200")


def strict() -> bool:
    return os.environ.get("SDAG_STRICT", "") == "1"


def note_fallback(where: str, reason: str, exc: BaseException = None, fatal_in_strict: bool = True) -> None:
    """``fatal_in_strict=False``: an alternative of (nearly) equal cost, recorded but never an error."""
    key = f"{where}: {reason}"
    first = key not in fallbacks
    fallbacks[key] = fallbacks.get(key, 0) + 1
    if strict() and fatal_in_strict:
        raise SdagB200Error(f"SDAG_STRICT=1 and the fast path was left at {key}") from exc
    if first:
        _log.warning("sdag_b200 fallback at %s%s", key, f" ({exc!r})" if exc is not None else "")


# kernels launched through this binding since import (bench.py reports it as "gpu_launches")
launch_count = 0


def count_launches(n: int) -> None:
    global launch_count
    launch_count += n
