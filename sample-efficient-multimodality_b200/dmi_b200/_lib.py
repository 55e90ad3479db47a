"""ctypes binding of ``libdmi_b200.so`` (C ABI declared in ``include/dmi_b200.h``).

There is deliberately no fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "..", "csrc", "libdmi_b200.so")

c_void_p, c_int, c_int64, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_float


class MlpArgs(C.Structure):
    """Mirror of ``struct dmi_mlp_args`` (include/dmi_b200.h)."""
    _fields_ = [
        ("B", c_int64), ("D", c_int64), ("H", c_int64), ("r", c_int64),
        ("flags", C.c_int32), ("grad_scale", c_float), ("dropout_p", c_float), ("_pad", C.c_int32),
        ("x", c_void_p), ("ldx", c_int64),
        ("dy", c_void_p), ("lddy", c_int64),
        ("keep", c_void_p),
        ("w1ext", c_void_p), ("w2ext", c_void_p), ("w2text", c_void_p),
        ("a0t", c_void_p), ("a1t", c_void_p), ("b0", c_void_p), ("b1", c_void_p),
        ("bias0", c_void_p), ("bias1", c_void_p),
        ("xext", c_void_p), ("pre", c_void_p), ("hext", c_void_p), ("dyext", c_void_p), ("dpre", c_void_p), ("du", c_void_p),
        ("y", c_void_p), ("ldy", c_int64),
        ("y_bf16", c_void_p), ("ldy_bf16", c_int64),
        ("dA0", c_void_p), ("dB0", c_void_p), ("dbeta0", c_void_p),
        ("dA1", c_void_p), ("dB1", c_void_p), ("dbeta1", c_void_p),
        ("dW1", c_void_p), ("db1", c_void_p), ("dW2", c_void_p), ("db2", c_void_p),
        ("ev_layer1_grads", c_void_p),
    ]


MLP_STOP_AFTER_FIRST_ACT = 1
MLP_NO_ADAPTER = 2
MLP_X_PREPACKED = 4
MLP_BASE_GRADS = 8
MLP_DROPOUT = 16

# name -> (restype, argtypes); every symbol of include/dmi_b200.h must be listed here (tests check it).
SIGNATURES = {
    "dmi_version": (c_int, []),
    "dmi_last_error": (C.c_char_p, []),
    "dmi_num_sms": (c_int, []),
    "dmi_launch_count": (c_int64, []),
    "dmi_set_option": (c_int, [C.c_char_p, c_int]),
    "dmi_gemm_mn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p, c_int64, c_int, c_void_p]),
    "dmi_gemm_tn": (c_int, [c_int, c_int, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_float, c_void_p,
                            c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "dmi_outer_reduce": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int,
                                 c_void_p, c_float, c_void_p]),
    "dmi_projector_pack_base": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "dmi_adapter_pack": (c_int, [c_void_p] * 8 + [c_int64, c_int64, c_int64, c_float] + [c_void_p] * 10),
    "dmi_merge_adapter": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_float,
                                  c_void_p, c_int64, c_void_p, c_void_p]),
    "dmi_adapted_mlp_fwd": (c_int, [C.POINTER(MlpArgs), c_void_p]),
    "dmi_adapted_mlp_bwd": (c_int, [C.POINTER(MlpArgs), c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.abspath(LIB_PATH)
    if not os.path.exists(path):
        raise RuntimeError(
            f"libdmi_b200.so not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C sample-efficient-multimodality_b200/csrc`). There is no CPU / PyTorch fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().dmi_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with status {rc}: {last_error()}")
