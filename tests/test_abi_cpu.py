"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dmi_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from dmi_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 9
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmi_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in dmi_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string_without_gpu():
    from dmi_b200 import _lib
    lib = _lib.load()
    assert lib.dmi_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_mlp_args_struct_layout_matches_header():
    """field order of the ctypes mirror == field order of struct dmi_mlp_args"""
    from dmi_b200._lib import MlpArgs
    src = open(HEADER).read()
    body = src[src.index("typedef struct dmi_mlp_args {"):src.index("} dmi_mlp_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            fields.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
    assert fields == [f[0] for f in MlpArgs._fields_]
    assert ctypes.sizeof(MlpArgs) % 8 == 0


def test_cpu_tensors_are_rejected_not_emulated():
    import torch
    from dmi_b200 import ops
    a = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm_tn(a, a, out0=torch.zeros(8, 8))
