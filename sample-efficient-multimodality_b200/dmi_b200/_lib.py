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
    "dmi_skinny_rows": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "dmi_outer_reduce": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int,
                                 c_void_p, c_float, c_void_p]),
    "dmi_panel_fused_tc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                   c_float, c_int64, c_int64, c_int64, c_void_p]),
    "dmi_panel_tc_project": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "dmi_panel_fused_tc32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                     c_int64, c_void_p, c_float, c_int64, c_int64, c_int64, c_void_p]),
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


# ---- part 2 of the ABI: augmentation, hypernetwork, splice -----------------------------------------------------------
MAX_GEN_LAYERS = 4
AUG_NORMALIZE = 1
_F4 = c_void_p * MAX_GEN_LAYERS


class AugmentArgs(C.Structure):
    """Mirror of ``struct dmi_augment_args``."""
    _fields_ = [
        ("B", c_int64), ("K", c_int64), ("D", c_int64), ("Dh", c_int64), ("D_src", c_int64),
        ("flags", C.c_int32), ("_pad", C.c_int32),
        ("mm", c_void_p), ("ld_mm", c_int64),
        ("sup", c_void_p), ("ld_sup", c_int64),
        ("txt", c_void_p), ("ld_txt", c_int64),
        ("prefix", c_void_p), ("R", c_void_p), ("perm", c_void_p), ("sign", c_void_p),
        ("mm_out", c_void_p), ("ld_mm_out", c_int64),
        ("mm_out_bf16", c_void_p), ("ld_mm_bf16", c_int64),
        ("z", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", C.c_uint64),
    ]


class HypernetArgs(C.Structure):
    """Mirror of ``struct dmi_hypernet_args``."""
    _fields_ = [
        ("S_z", c_int64), ("NQ", c_int64), ("D", c_int64), ("n_layers", c_int64),
        ("out_scale", c_float), ("dropout_p", c_float), ("overwrite_gen_grads", C.c_int32), ("_pad", C.c_int32),
        ("z", c_void_p), ("ldz", c_int64),
        ("prefix_tokens", c_void_p),
        ("pe", c_void_p), ("ldpe", c_int64),
        ("wq", c_void_p), ("bq", c_void_p), ("wk", c_void_p), ("bk", c_void_p), ("wv", c_void_p), ("bv", c_void_p),
        ("gen_w", _F4), ("gen_b", _F4), ("gen_out", c_int64 * MAX_GEN_LAYERS),
        ("keep", c_void_p),
        ("w_out", _F4),
        ("stash", c_void_p),
        ("dw", _F4),
        ("scratch", c_void_p),
        ("dprefix", c_void_p), ("dwq", c_void_p), ("dbq", c_void_p), ("dwk", c_void_p), ("dbk", c_void_p), ("dwv", c_void_p), ("dbv", c_void_p),
        ("dgen_w", _F4), ("dgen_b", _F4),
    ]


SIGNATURES.update({
    "dmi_l2_normalize": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_void_p]),
    "dmi_augment_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "dmi_augment": (c_int, [C.POINTER(AugmentArgs), c_void_p]),
    "dmi_hypernet_stash_floats": (c_int64, [c_int64, c_int64, c_int64]),
    "dmi_hypernet_scratch_floats": (c_int64, [c_int64, c_int64, c_int64]),
    "dmi_hypernet_stash_code_offset": (c_int64, [c_int64, c_int64, c_int64]),
    "dmi_hypernet_fwd": (c_int, [C.POINTER(HypernetArgs), c_void_p]),
    "dmi_hypernet_bwd": (c_int, [C.POINTER(HypernetArgs), c_void_p]),
    "dmi_hypernet_pool": (c_int, [C.POINTER(HypernetArgs), c_void_p, c_float, c_void_p]),
    "dmi_hypernet_generate": (c_int, [C.POINTER(HypernetArgs), c_void_p, c_void_p]),
    "dmi_haar_workspace_bytes": (c_int64, [c_int64]),
    "dmi_haar_orthogonal": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, C.c_uint64, c_void_p]),
    "dmi_gather_rows": (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int,
                                c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "dmi_splice": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int,
                           c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
})


# ---- part 3 of the ABI: fused clip-grad-norm + AdamW ------------------------------------------------------------------
class OptTensor(C.Structure):
    """Mirror of ``struct dmi_opt_tensor``."""
    _fields_ = [("p", c_void_p), ("g", c_void_p), ("m", c_void_p), ("v", c_void_p), ("n", c_int64)]


SIGNATURES.update({
    "dmi_grad_sqnorm": (c_int, [C.POINTER(OptTensor), c_int, c_void_p, c_void_p]),
    "dmi_grad_clip": (c_int, [C.POINTER(OptTensor), c_int, c_float, c_void_p, c_void_p]),
    "dmi_adamw_step": (c_int, [C.POINTER(OptTensor), c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_int64,
                               c_float, c_void_p, c_int, c_void_p]),
})


# ---- part 4 of the ABI: one-shot all-reduce over peer-mapped memory ----------------------------------------------------
SIGNATURES.update({
    "dmi_allreduce_flag_words": (c_int64, []),
    "dmi_allreduce_oneshot": (c_int, [C.POINTER(c_void_p), C.POINTER(c_void_p), c_void_p, c_int, c_int, c_void_p, c_int64, c_float, C.c_uint32, c_void_p]),
})
