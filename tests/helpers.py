"""Shared helpers for the parity tests."""
import glob
import os

import torch

from fabind_b200.synthetic import make_batch
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
    b = make_batch(embed=r["hidden"], **r["batch"])
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
