"""ctypes binding of libjaeger_b200.so (include/jaeger_b200.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded the
import raises, and without a B200 `Context()` raises (jg_ctx_create fails).
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_double, c_int16, c_int32, c_int64, c_uint8, c_uint16, c_uint32, c_void_p, c_float
from pathlib import Path

from .plan import HeadDesc, LayerDesc

LIB_PATH = Path(__file__).resolve().parent / "libjaeger_b200.so"


class JaegerB200Error(RuntimeError):
    pass


def _load():
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m jaeger_b200.build` "
            "(nvcc, sm_100a). The jaeger_b200 hot path has no CPU fallback.")
    return ctypes.CDLL(str(LIB_PATH))


lib = _load()

_P = c_void_p   # device pointers travel as integers (torch tensor.data_ptr())
_SIGNATURES = {
    "jg_last_error": (ctypes.c_char_p, []),
    "jg_version": (c_int32, []),
    "jg_ctx_create": (c_int32, [c_int32, POINTER(c_void_p)]),
    "jg_ctx_create_on_stream": (c_int32, [c_int32, c_void_p, POINTER(c_void_p)]),
    "jg_ctx_destroy": (c_int32, [c_void_p]),
    "jg_ctx_sync": (c_int32, [c_void_p]),
    "jg_ctx_stream": (c_void_p, [c_void_p]),
    "jg_ctx_launch_count": (c_int64, [c_void_p]),
    "jg_fasta_scan": (c_int32, [ctypes.c_char_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "jg_fasta_load": (c_int32, [ctypes.c_char_p, _P, POINTER(c_int64), ctypes.c_char_p]),
    "jg_fasta_open": (c_int32, [ctypes.c_char_p, c_int64, c_int64, POINTER(c_void_p)]),
    "jg_fasta_next": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_int64, _P, POINTER(c_int64), ctypes.c_char_p,
                                POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "jg_fasta_close": (c_int32, [c_void_p]),
    "jg_pack_bases": (c_int32, [c_void_p, _P, c_int64, _P, _P]),
    "jg_dust_mask": (c_int32, [c_void_p, _P, _P, _P, _P, _P, _P, c_int64, c_int32, _P]),
    "jg_plan_windows": (c_int32, [POINTER(c_int64), c_int64, c_int32, c_int32, c_int32, c_double, c_int32, c_int64,
                                  c_int32, POINTER(c_int64), POINTER(c_int32), POINTER(c_int64), POINTER(c_int32),
                                  POINTER(c_int32), POINTER(c_uint8)]),
    "jg_encode_windows": (c_int32, [c_void_p, _P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32,
                                    POINTER(c_uint8), c_int32, _P, _P, _P]),
    "jg_model_create": (c_int32, [c_void_p, POINTER(LayerDesc), c_int32, POINTER(HeadDesc), c_int32, c_int32,
                                  POINTER(c_void_p)]),
    "jg_model_destroy": (c_int32, [c_void_p]),
    "jg_model_max_windows": (c_int64, [c_void_p, c_int32, c_int64]),
    "jg_model_forward": (c_int32, [c_void_p, c_void_p, _P, _P, c_int64, c_int32, c_int32, _P, _P, _P, _P, c_int32]),
    "jg_model_workspace_bytes": (c_int64, [c_void_p]),
    "jg_model_flops_per_window": (c_double, [c_void_p, c_int32]),
    "jg_model_set_profiling": (c_int32, [c_void_p, c_int32]),
    "jg_model_kernel_names": (c_int32, [c_void_p, ctypes.c_char_p, c_int32]),
    "jg_model_get_profile": (c_int32, [c_void_p, c_int32, POINTER(c_double), POINTER(c_int64), POINTER(c_double)]),
    "jg_aggregate_contigs": (c_int32, [c_void_p, _P, _P, _P, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "jg_smooth_scores": (c_int32, [c_void_p, _P, _P, c_int64, c_int32, c_int32, _P]),
    "jg_segment_scores_batched": (c_int32, [c_void_p, _P, _P, c_int32, c_int64, c_int32, c_int32, _P, _P]),
    "jg_segment_scores": (c_int32, [c_void_p, _P, c_int32, c_int32, c_int32, _P, _P]),
    "jg_legacy_reliability": (c_int32, [c_void_p, _P, c_int64, c_int32, _P, _P, _P, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                        _P, c_int32, _P, _P]),
    "jg_sw_scan": (c_int32, [c_void_p, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P]),
    "jg_sw_trace": (c_int32, [c_void_p, _P, _P, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "jg_refine_contigs": (c_int32, [c_void_p, _P, _P, c_int64, c_int64, c_int32, _P, c_int32, c_int32, c_int32, c_double, _P, _P, _P, _P, _P]),
    "jg_viterbi_decode": (c_int32, [c_void_p, _P, _P, c_int32, c_int64, c_int32, _P, _P, _P]),
}
EXPORTED = tuple(_SIGNATURES)
for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here == header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int) -> None:
    if rc != 0:
        raise JaegerB200Error(lib.jg_last_error().decode("utf-8", "replace"))


class Context:
    """One CUDA context handle (device + stream) of the library."""

    def __init__(self, device: int = 0, stream: int | None = None):
        h = c_void_p()
        if stream:
            check(lib.jg_ctx_create_on_stream(int(device), c_void_p(int(stream)), ctypes.byref(h)))
        else:
            check(lib.jg_ctx_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)

    def sync(self) -> None:
        check(lib.jg_ctx_sync(self.handle))

    @property
    def stream(self) -> int:
        return int(lib.jg_ctx_stream(self.handle) or 0)

    @property
    def launch_count(self) -> int:
        return int(lib.jg_ctx_launch_count(self.handle))

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib.jg_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
