"""FABind+ L2 wrapper `FABindPlus` on the GPU (pocket stage -> radius head -> soft centre -> crop + re-centring -> FABind+
docking stack -> MLP distance head on the propagated pair embedding) through the C ABI, against goldens generated from the
unmodified reference and against the CPU oracle at the published width."""
import json
import os

import pytest
import torch

from oracle import fabind_plus_oracle_l2 as l2p
from oracle.det_weights import det_state_dict
from fabind_b200.config import published_args_plus
from fabind_b200.plus import FABindPlus
from fabind_b200.synthetic import make_docking_batch
from helpers import l2plus_golden_files, load_l2plus_golden, rel_err

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _compare(out, ref, tag, skip=()):
    rec = {}
    assert len(out) == len(ref)
    for i, (a, b) in enumerate(zip(out, ref)):
        if i in skip:
            continue
        if torch.is_tensor(b):
            a = a.cpu()
            assert tuple(a.shape) == tuple(b.shape), (i, a.shape, b.shape)
            if b.dtype.is_floating_point:
                rec[f"out{i}"] = rel_err(a, b)
            else:
                assert torch.equal(a.to(b.dtype), b), i
        else:
            assert a == b, i
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "parity.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=tag, **rec)) + "\n")
    return rec


@pytest.mark.parametrize("path", l2plus_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_l2plus_golden(path):
    g, r, args, data, sd = load_l2plus_golden(path)
    m = FABindPlus(args, r["emb"], r["pemb"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(g["shapes"])       # drop-in state_dict layout
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    d = data.to("cuda")
    out = m(d, stage=2)
    torch.cuda.synchronize()
    # element 11 = relu(radius head): when the head sits on the relu threshold (published-crop fixture: 0 / 2e-4 / 0) a
    # relative bound on the output is meaningless; its pre-activation is covered by the second fixture (radius ~ 8.8)
    rec = _compare(out, g["forward"], "l2plus_golden_forward", skip=(11,) if float(g["forward"][11].max()) < 1e-2 else ())
    assert max(rec.values()) < 1e-4, rec
    assert float((out[11].cpu() - g["forward"][11]).abs().max()) < 1e-4
    assert rel_err(d.coords.cpu(), g["coords_after"]) < 1e-5          # in-place shift of data.coords (model.py:257)
    inf = m.inference(data.to("cuda"))
    assert rel_err(inf[0].cpu(), g["inference"]) < 1e-4
    # stage 1: the dataloader's pocket (main_fabind.py:179 evaluates the test set this way), incl. the in-place coordinate shifts
    d1 = data.to("cuda")
    out1 = m(d1, stage=1)
    torch.cuda.synchronize()
    rec1 = _compare(out1, g["forward_stage1"], "l2plus_golden_forward_stage1", skip=(11,) if float(g["forward_stage1"][11].max()) < 1e-2 else ())
    assert max(rec1.values()) < 1e-4, rec1
    assert rel_err(d1.coords.cpu(), g["coords_after_stage1"]) < 1e-5
    assert rel_err(d1['complex'].node_coords.cpu(), g["complex_coords_after_stage1"]) < 1e-5


def test_l2plus_vs_oracle_published_width():
    """hidden 512 / pocket 128, 5 layers x 2 iterations, whole proteins of 150-300 residues"""
    args = published_args_plus(mean_layers=5, n_iter=2)
    m = FABindPlus(args, 512, 128)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 61)
    m.load_state_dict(sd, strict=True)
    data = make_docking_batch(2, seed=11, n_c_range=(10, 30), L_range=(150, 300))
    with torch.no_grad():
        ref = l2p.forward_stage2(sd, args, data.clone())
    m = m.cuda().eval()
    out = m(data.to("cuda"), stage=2)
    torch.cuda.synchronize()
    rec = _compare(out, ref, "l2plus_oracle_published_width", skip=(11,))
    assert max(rec.values()) < 1e-4, rec
    assert float((out[11].cpu() - ref[11]).abs().max()) < 1e-3


def test_l2plus_bf16_runs_and_is_close():
    args = published_args_plus(mean_layers=2, n_iter=2)
    m = FABindPlus(args, 256, 128)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 62)
    g = torch.Generator().manual_seed(3)
    for k in sd:      # reference-style small coordinate heads (xavier gain 0.001): bf16 deviations stay out of the edge sets
        if k.endswith("coord_mlp.linear2.weight"):
            sd[k] = (torch.rand(sd[k].shape, generator=g) * 2 - 1) * 0.001 * (6.0 / (sd[k].shape[1] + 1)) ** 0.5
    m.load_state_dict(sd, strict=True)
    data = make_docking_batch(2, seed=12, n_c_range=(10, 30), L_range=(150, 300))
    m = m.cuda().eval()
    a = m(data.to("cuda"), stage=2)
    m.precision = "bf16"
    b = m(data.to("cuda"), stage=2)
    torch.cuda.synchronize()
    assert a[2].shape == b[2].shape
    assert float((a[0] - b[0]).abs().max()) < 0.5                      # predicted ligand coordinates, Angstrom
    assert float((a[2] - b[2]).abs().mean()) < 0.15                    # distance map head, Angstrom (range 0-15)
