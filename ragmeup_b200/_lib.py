"""ctypes binding of libragmeup_b200.so — the ONLY way the Python host code reaches the GPU.

There is no CPU fallback: if the library is missing or no CUDA device is present every product
call raises.  (``include/ragmeup_b200.h`` is the authoritative declaration of these symbols.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libragmeup_b200.so")

_lock = threading.Lock()
_lib = None


class RmuError(RuntimeError):
    pass


class BertConfigC(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("ffn", C.c_int32), ("max_pos", C.c_int32), ("type_vocab", C.c_int32), ("num_labels", C.c_int32),
                ("ln_eps", C.c_float)]


# name -> (restype, argtypes); mirrors include/ragmeup_b200.h
SIGNATURES = {
    "rmu_last_error": (C.c_char_p, []),
    "rmu_version": (C.c_int, []),
    "rmu_launch_count": (C.c_uint64, []),
    "rmu_profile_enable": (None, [C.c_int]),
    "rmu_profile_reset": (None, []),
    "rmu_profile_read": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "rmu_index_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "rmu_index_destroy": (None, [C.c_void_p]),
    "rmu_index_reserve": (C.c_int, [C.c_void_p, C.c_int64]),
    "rmu_index_add": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "rmu_index_set_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "rmu_index_size": (C.c_int64, [C.c_void_p]),
    "rmu_index_dim": (C.c_int, [C.c_void_p]),
    "rmu_index_metric": (C.c_int, [C.c_void_p]),
    "rmu_index_clear": (C.c_int, [C.c_void_p]),
    "rmu_index_data": (C.c_void_p, [C.c_void_p]),
    "rmu_index_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "rmu_index_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "rmu_debug_scan_tile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "rmu_index_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "rmu_topk_merge": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "rmu_topk_merge_strided": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "rmu_adjacent_cosine_distance": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "rmu_mmr_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "rmu_bm25_create": (C.c_int, [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_double, C.POINTER(C.c_void_p)]),
    "rmu_bm25_destroy": (C.c_int, [C.c_void_p]),
    "rmu_bm25_size": (C.c_int64, [C.c_void_p]),
    "rmu_bm25_terms": (C.c_int64, [C.c_void_p]),
    "rmu_bm25_search": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "rmu_bm25_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "rmu_bm25_csr_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "rmu_bm25_csr_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "rmu_bm25_csr_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "rmu_bm25_csr_free": (C.c_int, [C.c_void_p]),
    "rmu_encoder_create": (C.c_int, [C.POINTER(BertConfigC), C.POINTER(C.c_void_p), C.c_int, C.c_int,
                                     C.POINTER(C.c_void_p)]),
    "rmu_encoder_destroy": (None, [C.c_void_p]),
    "rmu_encoder_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p]),
    "rmu_encoder_classify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p]),
    "rmu_encoder_hidden": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "rmu_encoder_embed_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rmu_encoder_classify_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]),
}


def lib() -> C.CDLL:
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            # the library is built in-tree and normally travels with the checkout; build it when absent
            try:
                from . import build as _build
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RmuError(f"{LIB_PATH} is not built and building it failed ({e}): run "
                               f"`python -m ragmeup_b200.build` (there is no CPU fallback for this path)") from e
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rmu_last_error()
        raise RmuError(f"{what or 'ragmeup_b200'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(lib().rmu_launch_count())


def require_cuda():
    """The product path is CUDA-only (the reference's force_cpu / device='cpu' has no equivalent)."""
    import torch
    if not torch.cuda.is_available():
        raise RmuError("ragmeup_b200 needs a CUDA device (sm_100a); there is no CPU path")
    return torch


def stream_ptr(torch_stream=None) -> int:
    import torch
    s = torch_stream if torch_stream is not None else torch.cuda.current_stream()
    return int(s.cuda_stream)


PROF_CLASSES = ["scan", "finalize", "exact", "merge", "gemm", "attention", "layernorm", "embedding", "pool_head", "misc",
                "scan_pass2"]


def profile_enable(on: bool) -> None:
    lib().rmu_profile_enable(int(bool(on)))


def profile_reset() -> None:
    lib().rmu_profile_reset()


def profile_read() -> dict:
    """{class: (total_ms, launches)} since the last reset (synchronises the recorded events)."""
    out = {}
    for i, name in enumerate(PROF_CLASSES):
        ms, n = C.c_double(0), C.c_int64(0)
        check(lib().rmu_profile_read(i, C.byref(ms), C.byref(n)), "rmu_profile_read")
        out[name] = (ms.value, int(n.value))
    return out
