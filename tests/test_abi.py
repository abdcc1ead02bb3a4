"""The C-ABI library loads on a CPU-only machine and exports every symbol include/fabind_b200.h declares."""
import os
import re

from fabind_b200 import _lib
from fabind_b200.weights import slots

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "fabind_b200.h")).read()
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    l = _lib.lib()
    for name in declared:
        assert hasattr(l, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert l.fb_abi_version() == _lib.ABI_VERSION == 6


def test_weight_slots_are_disjoint_and_cover_arena():
    l = _lib.lib()
    for flavour in (0, 1):
        for hidden, L in [(128, 1), (512, 4), (64, 2)]:
            s = slots(hidden, L, flavour)
            end = 0
            for name, r, c, off in s:
                assert off >= end, name
                end = off + r * c
            assert end <= l.fb_weight_arena_elems_f(hidden, L, flavour)


def _header_prototypes():
    """{name: [parameter class]} parsed from include/fabind_b200.h; classes: 'ptr', 'i32', 'i64', 'f32', 'u32'"""
    hdr = open(os.path.join(ROOT, "include", "fabind_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int32_t|int64_t|void\*|const void\*|const int32_t\*)\s+(fb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        kinds = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                if "*" in p:
                    kinds.append("ptr")
                elif re.match(r"(const )?int32_t\b", p):
                    kinds.append("i32")
                elif re.match(r"(const )?int64_t\b", p):
                    kinds.append("i64")
                elif re.match(r"(const )?uint32_t\b", p):
                    kinds.append("u32")
                elif re.match(r"(const )?float\b", p):
                    kinds.append("f32")
                else:
                    raise AssertionError(f"{name}: unparsed parameter '{p}'")
        protos[name] = kinds
    return protos


def test_ctypes_prototypes_match_the_header():
    """every binding in fabind_b200/_lib.py has the header's parameter count and, position by position, the header's kind of
    parameter (pointer / int32 / int64 / float): a transposed or missing argument in a binding would corrupt a launch silently"""
    import ctypes as C
    protos = _header_prototypes()
    assert set(protos) == set(_lib.EXPORTS), set(protos) ^ set(_lib.EXPORTS)

    def kind(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        return {C.c_int32: "i32", C.c_int64: "i64", C.c_float: "f32", C.c_uint32: "u32"}[t]
    for name, (res, args) in _lib.EXPORTS.items():
        got = [kind(t) for t in args]
        assert got == protos[name], (name, got, protos[name])
