"""Training-path groundwork (BASELINE config 5, later rounds): the oracles' AUTOGRAD path against parameter gradients of the
unmodified reference (scripts/make_golden.py::main_grad): eval mode (no dropout), refine='refine_coord' semantics (only the last
refinement iteration carries gradients, att_model.py:227-245 / P:196-218), fixed linear read-out of (X, H[, pair]).  Parameters
the reference never uses (`att_i.inter_layer.*`, SURVEY 8e: find_unused_parameters) have no gradient on either side."""
import glob
import os

import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import fabind_plus_oracle as orcp
from helpers import GOLDEN_DIR, load_golden, rel_err


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_*.pt"))), ids=lambda p: p.split("/")[-1][:-3])
def test_oracle_gradients_match_reference(path):
    g, r, b, sd, cfg = load_golden(path)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    gen = torch.Generator().manual_seed(r["readout_seed"])
    f = orcp.model_forward if r["flavour"] == "plus" else orc.model_forward
    out = f(sd, cfg, b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, b.LAS_edge_index, b.X_LAS,
            grad_last_iter_only=True)
    rx, rh = torch.randn(out[0].shape, generator=gen), torch.randn(out[1].shape, generator=gen)
    loss = (out[0] * rx).sum() + (out[1] * rh).sum()
    if r["flavour"] == "plus":
        loss = loss + (out[2] * (torch.randn(out[2].shape, generator=gen) * 0.1)).sum()
    loss.backward()
    assert abs(float(loss) - g["loss"]) < 1e-4 * abs(g["loss"])
    n_checked = 0
    gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
    for k, ref in g["grads"].items():
        mine = sd[k].grad
        if ref is None:
            assert mine is None or float(mine.abs().max()) == 0.0, k            # unused parameter on both sides
            continue
        assert mine is not None, k
        # fp32 accumulation noise is absolute: tensors whose gradients are 1e-5 of the largest ones get an absolute floor
        err = float((mine - ref).abs().max())
        assert err < 2e-4 * float(ref.abs().max()) + 2e-7 * gmax, (k, err, float(ref.abs().max()), gmax)
        n_checked += 1
    assert n_checked >= 80
