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


def header_prototypes():
    """{name: (return kind, [argument kinds])} of every function prototype in the header; kinds: ptr / i64 / i32 / f32 / f64"""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", src, flags=re.S)
    src = re.sub(r"typedef enum \w* *\{.*?\} \w+;", "", src, flags=re.S)

    def kind(decl):
        decl = " ".join(decl.split())
        if "*" in decl:
            return "ptr"
        base = decl.split()
        for t, k in (("int64_t", "i64"), ("uint64_t", "i64"), ("size_t", "i64"), ("float", "f32"), ("double", "f64"), ("int", "i32"), ("int32_t", "i32"), ("uint32_t", "i32")):
            if t in base:
                return k
        raise AssertionError(f"unknown C type in header prototype: {decl!r}")

    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(dmi_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argk = [] if args in ("", "void") else [kind(a) for a in args.split(",")]
        out[name] = (kind(ret + " x") if ret != "void" else "void", argk)
    return out


def test_ctypes_signatures_match_header_prototypes():
    """argument count and kind (pointer / 64-bit / 32-bit integer / float) of every ctypes signature == the prototype in the header,
    so that a new entry point cannot be bound with a shifted or mistyped argument list"""
    from dmi_b200 import _lib
    protos = header_prototypes()
    assert sorted(protos) == declared_symbols()

    def ckind(t):
        if t is None:
            return "void"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or isinstance(t, type) and issubclass(t, ctypes._Pointer):
            return "ptr"
        if t in (ctypes.c_int64, ctypes.c_uint64, ctypes.c_size_t):
            return "i64"
        if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_uint32):
            return "i32"
        if t is ctypes.c_float:
            return "f32"
        if t is ctypes.c_double:
            return "f64"
        raise AssertionError(f"unknown ctypes type {t}")

    for name, (ret, args) in protos.items():
        restype, argtypes = _lib.SIGNATURES[name]
        assert [ckind(t) for t in argtypes] == args, name
        assert ckind(restype) == ret, name


def header_struct_fields(name):
    """[(field, c_type_text, array_len)] of `typedef struct <name> {...}` in declaration order"""
    src = open(HEADER).read()
    body = src[src.index("typedef struct %s {" % name):src.index("} %s;" % name)]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        first, *rest = decl.split(",")
        m = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[[A-Za-z0-9_]+\])?$", first.strip())
        base = m.group(1).strip()
        base_type = base.rstrip("* ").strip()
        for part in [first] + rest:
            mm = re.match(r"(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[[A-Za-z0-9_]+\])?$", part.strip())
            is_ptr = "*" in mm.group(1) or (part is first and "*" in base)
            out.append((mm.group(2), "ptr" if is_ptr else base_type, mm.group(3)))
    return out


@pytest.mark.parametrize("cname,pyname", [("dmi_mlp_args", "MlpArgs"), ("dmi_augment_args", "AugmentArgs"), ("dmi_hypernet_args", "HypernetArgs"),
                                          ("dmi_opt_tensor", "OptTensor")])
def test_ctypes_struct_layout_matches_header(cname, pyname):
    """field order, pointer-ness, scalar width and array length of each ctypes mirror == the C struct in the header"""
    from dmi_b200 import _lib
    cls = getattr(_lib, pyname)
    hdr = header_struct_fields(cname)
    assert [h[0] for h in hdr] == [f[0] for f in cls._fields_]
    width = {"int64_t": 8, "uint64_t": 8, "int32_t": 4, "float": 4}
    for (name, ctype_txt, arr), (pname, ptype) in zip(hdr, cls._fields_):
        n = 1
        if arr is not None:
            n = _lib.MAX_GEN_LAYERS if not arr.strip("[]").isdigit() else int(arr.strip("[]"))
        expect = (8 if ctype_txt == "ptr" else width[ctype_txt]) * n
        assert ctypes.sizeof(ptype) == expect, (cname, name, ctype_txt, arr)
    assert ctypes.sizeof(cls) % 8 == 0


def test_cpu_tensors_are_rejected_not_emulated():
    import torch
    from dmi_b200 import ops
    a = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm_tn(a, a, out0=torch.zeros(8, 8))
