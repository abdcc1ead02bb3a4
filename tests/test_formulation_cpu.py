"""The refactored formulation that the CUDA launch sequence implements (tests/emulate_packed.py mirrors
csrc/forward.cu with the same packed weight arena and node order) against the golden vectors of the
unmodified reference.  Validates fabind_b200/weights.py and fabind_b200/layout.py without a GPU."""
import pytest
import torch

from helpers import golden_files, load_golden, rel_err
from emulate_packed import forward_emulated


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_refactored_formulation_matches_reference(path):
    g, r, b, sd, cfg = load_golden(path)
    with torch.no_grad():
        X, H, stats = forward_emulated(sd, cfg, b)
    assert stats == [int(e[1].shape[1]) for e in g["edges"]]
    assert rel_err(X, g["X"]) < 1e-5
    assert rel_err(H, g["H"]) < 1e-4
