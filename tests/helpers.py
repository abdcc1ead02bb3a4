"""Shared helpers for the parity tests."""
import glob
import os

import torch

from fabind_b200.synthetic import make_batch, batch_from_recipe
from oracle.det_weights import det_state_dict
from oracle import fabind_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "v1_*.pt")))


def plus_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "plus_*.pt")))


def l2_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "l2_*.pt")))


def load_l2_golden(path):
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_docking_batch
    g = torch.load(path, map_location="cpu", weights_only=False)
    r = g["recipe"]
    args = published_args(mean_layers=r["mean_layers"], n_iter=r["n_iter"])
    data = make_docking_batch(**r["batch"])
    sd = det_state_dict(g["shapes"], r["weight_seed"])
    return g, r, args, data, sd


def load_golden(path):
    g = torch.load(path, map_location="cpu", weights_only=False)
    r = g["recipe"]
    b = batch_from_recipe(r["hidden"], r["batch"], GOLDEN_DIR)
    if r["far_ligand"]:
        nc = b.n_c[0]
        b.X[1:nc + 1] += 20.0
    sd = det_state_dict(g["shapes"], r["weight_seed"])
    cfg = orc.make_cfg(n_layers=r["n_layers"], n_iter=r["n_iter"])
    return g, r, b, sd, cfg


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): the '1e-4 rel' metric of the north star, taken against the
    tensor's own scale so that exact zeros do not blow it up."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2plus_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "l2plus_*.pt")))


def load_l2plus_golden(path):
    from fabind_b200.config import published_args_plus
    from fabind_b200.synthetic import make_docking_batch
    g = torch.load(path, map_location="cpu", weights_only=False)
    r = g["recipe"]
    args = published_args_plus(mean_layers=r["mean_layers"], n_iter=r["n_iter"], **r["args_over"])
    data = make_docking_batch(**r["batch"])
    sd = det_state_dict(g["shapes"], r["weight_seed"])
    if r["radius_bias"] is not None:
        sd["pocket_radius_head.linear2.bias"] = torch.full_like(sd["pocket_radius_head.linear2.bias"], r["radius_bias"])
    return g, r, args, data, sd


def compare_tuple(mine, ref, tol=1e-5):
    """element-wise comparison of a model.forward return tuple: exact for bool/int tensors and python scalars"""
    assert len(mine) == len(ref)
    for i, (a, b) in enumerate(zip(mine, ref)):
        if torch.is_tensor(b):
            assert tuple(a.shape) == tuple(b.shape), (i, a.shape, b.shape)
            if b.dtype in (torch.bool, torch.int32, torch.int64):
                assert torch.equal(a.cpu().to(b.dtype), b), i
            else:
                assert rel_err(a, b) < tol, (i, rel_err(a, b))
        else:
            assert a == b, i
