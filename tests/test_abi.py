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
    assert l.fb_abi_version() == _lib.ABI_VERSION == 3


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
