"""L2 wrapper (models/model.py): the CPU oracle against reference-generated goldens (anywhere) and against the
live reference (dev container); parameter names/shapes of the drop-in module against the reference's."""
import pytest
import torch

from oracle import fabind_oracle_l2 as l2, ref_shims
from helpers import l2_golden_files, load_l2_golden, rel_err


@pytest.mark.parametrize("path", l2_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_l2_oracle_matches_golden(path):
    g, r, args, data, sd = load_l2_golden(path)
    with torch.no_grad():
        out = l2.forward_stage2(sd, args, data.clone())
        out1 = l2.forward_eval(sd, args, data.clone(), 1)
        inf = l2.inference(sd, args, data.clone())
    for i, (a, b) in enumerate(list(zip(out, g["forward"])) + list(zip(out1, g["forward_stage1"]))):
        if torch.is_tensor(b):
            assert a.shape == b.shape, i
            if b.dtype in (torch.bool, torch.int32, torch.int64):
                assert torch.equal(a.to(b.dtype), b), i
            else:
                assert rel_err(a, b) < 1e-5, (i, rel_err(a, b))
        else:
            assert a == b, i
    assert rel_err(inf[0], g["inference"]) < 1e-5


@pytest.mark.parametrize("path", l2_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_l2_module_state_dict_matches_reference(path):
    from fabind_b200.model import IaBNet_mean_and_pocket_prediction_cls_coords_dependent as Net
    g, r, args, data, sd = load_l2_golden(path)
    m = Net(args, r["emb"], r["pemb"])
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == g["shapes"]
    m.load_state_dict(sd, strict=True)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree absent")
def test_l2_oracle_vs_live_reference():
    from oracle.det_weights import det_state_dict
    from fabind_b200.synthetic import make_docking_batch
    mods = ref_shims.load_reference_model_module()
    args = ref_shims.published_args(mean_layers=1, n_iter=2, gs_hard=True)
    m = mods.model.IaBNet_mean_and_pocket_prediction_cls_coords_dependent(args, 48, 32).eval()
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 43)
    m.load_state_dict(sd, strict=True)
    d = make_docking_batch(2, seed=5, L_range=(120, 200))
    with torch.no_grad():
        ref = m(d.clone(), stage=2)
        mine = l2.forward_stage2(sd, args, d.clone())
        ri, mi = m.inference(d.clone()), l2.inference(sd, args, d.clone())
    for a, b in zip(mine, ref):
        if torch.is_tensor(b) and b.dtype.is_floating_point:
            assert rel_err(a, b) < 1e-5
    assert rel_err(mi[0], ri[0]) < 1e-5


# ---- FABind+ wrapper (FABind_plus/fabind/models/model.py::FABindPlus) -------------------------------------------------
from oracle import fabind_plus_oracle_l2 as l2p          # noqa: E402
from helpers import l2plus_golden_files, load_l2plus_golden, compare_tuple   # noqa: E402


@pytest.mark.parametrize("path", l2plus_golden_files(), ids=lambda p: p.split("/")[-1][:-3])
def test_l2plus_oracle_matches_golden(path):
    g, r, args, data, sd = load_l2plus_golden(path)
    with torch.no_grad():
        d2 = data.clone()
        out = l2p.forward_stage2(sd, args, d2)
        inf = l2p.inference(sd, args, data.clone())
    compare_tuple(out, g["forward"])
    assert rel_err(d2.coords, g["coords_after"]) < 1e-6          # in-place shift of the ground-truth pose (model.py:257)
    assert rel_err(inf[0], g["inference"]) < 1e-5
    with torch.no_grad():                                        # stage 1: the dataloader's pocket (model.py:169-197)
        d1 = data.clone()
        out1 = l2p.forward_eval(sd, args, d1, 1)
    compare_tuple(out1, g["forward_stage1"])
    assert rel_err(d1.coords, g["coords_after_stage1"]) < 1e-6
    assert rel_err(d1['complex'].node_coords, g["complex_coords_after_stage1"]) < 1e-6


def test_l2plus_golden_present():
    assert len(l2plus_golden_files()) >= 2
