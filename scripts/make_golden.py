"""Generate tests/golden/*.pt by running the UNMODIFIED reference (dev container only).

    python scripts/make_golden.py

Each fixture stores the recipe (synthetic-batch arguments, deterministic-weight seed, key shapes)
and the reference's outputs; inputs and weights are regenerated from the recipe by the tests
(`fabind_b200.synthetic.make_batch`, `oracle.det_weights.det_state_dict`), so fixtures stay small.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims                      # noqa: E402
from oracle.det_weights import det_state_dict     # noqa: E402
from fabind_b200.synthetic import make_batch, batch_from_recipe      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (hidden, n_layers, n_iter, make_batch kwargs, weight seed, far_ligand)
    "v1_h64_l2_it3_ragged": (64, 2, 3, dict(n_complexes=3, seed=1, n_c_range=(8, 30), n_p_range=(40, 90)), 11, False),
    "v1_h128_l1_it1_cfg1": (128, 1, 1, dict(n_complexes=1, seed=0, n_c=30, n_p=200), 12, False),
    "v1_h32_l1_it2_noedge": (32, 1, 2, dict(n_complexes=2, seed=3, n_c=6, n_p=30), 13, True),
    "v1_h32_l1_it2_fallback": (32, 1, 2, dict(n_complexes=1, seed=4, n_c=5, n_p=24), 14, True),
    # real pockets (scripts/make_real_geometry.py; SURVEY.md §8d geometry option A), true pose randomly rotated
    "v1_h32_l2_it3_real": (32, 2, 3, dict(seed=2, geometry="real_geometry.npz", ids=["6efk", "6g3c", "6n93", "6npi"]), 15, False),
}


def build_reference(mods, hidden, n_layers, n_iter, wseed):
    args = ref_shims.published_args()
    scale = args.coordinate_scale
    m = mods.att_model.EfficientMCAttModel(
        args, hidden, hidden, 1, n_edge_feats=0, n_layers=n_layers, n_iter=n_iter,
        inter_cutoff=args.inter_cutoff, intra_cutoff=args.intra_cutoff,
        normalize_coord=lambda x: x / scale, unnormalize_coord=lambda x: x * scale).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
    return m, shapes


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = ref_shims.load_reference("v1")
    for name, (hidden, L, IT, bkw, wseed, far) in CASES.items():
        m, shapes = build_reference(mods, hidden, L, IT, wseed)
        b = batch_from_recipe(hidden, bkw, OUT)
        if far:  # push the first complex's ligand 100 A away: no inter edge -> fallback (att_model.py:85)
            nc = b.n_c[0]
            b.X[1:nc + 1] += 20.0
        traces = []
        hooks = []
        gnn = m.gnn
        last = {"it": -1}

        def pre_hook(mod, inp):
            last["it"] += 1
        hooks.append(gnn.register_forward_pre_hook(pre_hook))
        for i in range(L):
            for kind in ("gcl", "att"):
                def hk(mod, inp, out, tag=f"{kind}_{i}"):
                    if last["it"] == IT - 1:
                        traces.append((tag, out[0].detach().clone(), out[1].detach().clone()))
                hooks.append(getattr(gnn, f"{kind}_{i}").register_forward_hook(hk))
        edges = []
        orig = m.extract_edges.forward

        def rec(X, bid, seg, glb):
            r = orig(X, bid, seg, glb)
            edges.append((r[0].to(torch.int32).clone(), r[1].to(torch.int32).clone()))
            return r
        m.extract_edges.forward = rec
        with torch.no_grad():
            bb = b.clone()
            X, H = m(**bb.forward_args())
        for h in hooks:
            h.remove()
        torch.save({
            "recipe": dict(hidden=hidden, n_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed,
                           far_ligand=far, flavour="v1"),
            "shapes": shapes,
            "X": X.clone(), "H": H.clone(),
            "edges": edges,
            "trace_last_iter": traces,
            "torch": torch.__version__,
        }, os.path.join(OUT, name + ".pt"))
        print(name, "X", tuple(X.shape), "moved", float((X - b.X).abs().max()),
              "E_int", [int(e[1].shape[1]) for e in edges])


PLUS_CASES = {
    "plus_h64_l2_it2_ragged": (64, 2, 2, dict(n_complexes=3, seed=1, n_c_range=(8, 30), n_p_range=(40, 90)), 31, False),
    "plus_h128_l1_it1_cfg1": (128, 1, 1, dict(n_complexes=1, seed=0, n_c=30, n_p=200), 32, False),
    "plus_h32_l1_it2_fallback": (32, 1, 2, dict(n_complexes=1, seed=4, n_c=5, n_p=24), 34, True),
    "plus_h32_l2_it2_real": (32, 2, 2, dict(seed=3, geometry="real_geometry.npz", ids=["6g3c", "6npi"]), 35, False),
}


def main_plus():
    """FABind+ weight layout (LayerNorm MLPs, propagated pair embedding): EfficientMCAttModel.forward -> (X, H, pair)"""
    mods = ref_shims.load_reference("plus")
    for name, (hidden, L, IT, bkw, wseed, far) in PLUS_CASES.items():
        args = ref_shims.published_args_plus()
        scale = args.coordinate_scale
        m = mods.att_model.EfficientMCAttModel(
            args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, n_iter=IT, inter_cutoff=args.inter_cutoff,
            intra_cutoff=args.intra_cutoff, normalize_coord=lambda x: x / scale, unnormalize_coord=lambda x: x * scale).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
        b = batch_from_recipe(hidden, bkw, OUT)
        if far:
            nc = b.n_c[0]
            b.X[1:nc + 1] += 20.0
        traces, hooks, last = [], [], {"it": -1}
        hooks.append(m.gnn.register_forward_pre_hook(lambda mod, inp: last.__setitem__("it", last["it"] + 1)))
        for i in range(L):
            for kind in ("gcl", "att"):
                def hk(mod, inp, out, tag=f"{kind}_{i}"):
                    if last["it"] == IT - 1:
                        traces.append((tag, out[0].detach().clone(), out[1].detach().clone()))
                hooks.append(getattr(m.gnn, f"{kind}_{i}").register_forward_hook(hk))
        edges = []
        orig = m.extract_edges.forward

        def rec(X, bid, seg, glb):
            r = orig(X, bid, seg, glb)
            edges.append((r[0].to(torch.int32).clone(), r[1].to(torch.int32).clone()))
            return r
        m.extract_edges.forward = rec
        with torch.no_grad():
            X, H, pair = m(**b.clone().forward_args())
        for h in hooks:
            h.remove()
        torch.save({
            "recipe": dict(hidden=hidden, n_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, far_ligand=far, flavour="plus"),
            "shapes": shapes, "X": X.clone(), "H": H.clone(), "pair": pair.clone(), "edges": edges,
            "trace_last_iter": traces, "torch": torch.__version__,
        }, os.path.join(OUT, name + ".pt"))
        print(name, "X", tuple(X.shape), "pair", tuple(pair.shape), "moved", float((X - b.X).abs().max()),
              "E_int", [int(e[1].shape[1]) for e in edges])


def main_l2():
    """goldens for the L2 wrapper (models/model.py): forward(stage=2) in eval mode and inference()"""
    from fabind_b200.synthetic import make_docking_batch
    mods = ref_shims.load_reference_model_module()
    cases = {"l2_h64_p32_l2_it2": (64, 32, 2, 2, dict(n_complexes=3, seed=1), 41)}
    for name, (emb, pemb, L, IT, bkw, wseed) in cases.items():
        args = ref_shims.published_args(mean_layers=L, n_iter=IT)
        m = mods.model.IaBNet_mean_and_pocket_prediction_cls_coords_dependent(args, emb, pemb).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
        d = make_docking_batch(**bkw)
        with torch.no_grad():
            fwd = m(d.clone(), stage=2)
            fwd1 = m(d.clone(), stage=1)
            inf = m.inference(d.clone())
        torch.save({"recipe": dict(emb=emb, pemb=pemb, mean_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed),
                    "shapes": shapes, "forward": [t.clone() if torch.is_tensor(t) else t for t in fwd],
                    "forward_stage1": [t.clone() if torch.is_tensor(t) else t for t in fwd1],
                    "inference": inf[0].clone(), "torch": torch.__version__}, os.path.join(OUT, name + ".pt"))
        print(name, [tuple(t.shape) if torch.is_tensor(t) else t for t in fwd])


def patch_reference_dropout(model, L, seed, p, prefix=""):
    """Replace every nn.Dropout of the (unmodified) FABind+ reference model by a deterministic COLUMN-ONLY mask drawn from the
    library's mask function (fabind_b200/dropout.py): the mask depends on (seed + iteration, site, feature column) only, so it
    is invariant to the reference's row orders / dense padding and pins the placement and scaling of every mask."""
    import re
    import torch.nn as nn
    from fabind_b200.dropout import keep_mask, site_id, iter_seed
    state = dict(it=-1, stack_calls=0)

    class ColDrop(nn.Module):
        def __init__(self, resolve):
            super().__init__()
            self.resolve = resolve

        def forward(self, x):
            if not self.training:
                return x
            site = self.resolve()
            m = keep_mask(iter_seed(seed, state["it"]), site, 1, x.shape[-1], p, colonly=True)[0]
            return x * m

    def pre_hook(mod, inp):
        state["it"] += 1
        state["stack_calls"] = 0
    stack = model
    for q in [q for q in prefix.split(".") if q]:
        stack = getattr(stack, q)
    stack.gnn.register_forward_pre_hook(pre_hook)

    def stack_site():
        state["stack_calls"] += 1
        return site_id(-1, "stack_in" if state["stack_calls"] == 1 else "stack_out")
    table = {"edge_mlp.dropout1": "edge1", "edge_mlp.dropout2": "edge2", "node_mlp.dropout1": "node1", "node_mlp.dropout2": "node2",
             "cross_attn_module.p_attention_block.dropout": "patt", "cross_attn_module.c_attention_block.dropout": "catt",
             "cross_attn_module.p_transition.dropout1": "ptr1", "cross_attn_module.p_transition.dropout2": "ptr2",
             "cross_attn_module.c_transition.dropout1": "ctr1", "cross_attn_module.c_transition.dropout2": "ctr2",
             "cross_attn_module.pair_transition.dropout1": "pair1", "cross_attn_module.pair_transition.dropout2": "pair2"}
    n = 0
    for name, mod in list(stack.named_modules()):
        if not isinstance(mod, nn.Dropout):
            continue
        parent = stack
        parts = name.split(".")
        for q in parts[:-1]:
            parent = getattr(parent, q)
        if name == "gnn.dropout":
            new = ColDrop(stack_site)
        else:
            m = re.match(r"gnn\.(gcl_(\d+)|att_(\d+)|out_layer)\.(.*)", name)
            assert m, name
            layer = L if m.group(1) == "out_layer" else int(m.group(2) or m.group(3))
            rest = m.group(4)
            if m.group(1).startswith("att") and rest == "dropout":
                key = "agg"
            elif rest == "coord_mlp.dropout":
                key = "acoord" if m.group(1).startswith("att") else "gcoord"
            else:
                key = table[rest]
            new = ColDrop(lambda s=site_id(layer, key): s)
        setattr(parent, parts[-1], new)
        n += 1
    return n


def main_plus_dropout():
    """FABind+ stack in train() mode (the reference's sampling mode) with every nn.Dropout replaced by a column-only mask"""
    mods = ref_shims.load_reference("plus")
    cases = {"plusdrop_h64_l2_it2": (64, 2, 2, dict(n_complexes=2, seed=1, n_c_range=(8, 20), n_p_range=(40, 70)), 35, 1234, 0.1)}
    for name, (hidden, L, IT, bkw, wseed, dseed, pdrop) in cases.items():
        args = ref_shims.published_args_plus(dropout=pdrop, random_n_iter=False)
        scale = args.coordinate_scale
        m = mods.att_model.EfficientMCAttModel(
            args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, n_iter=IT, inter_cutoff=args.inter_cutoff,
            intra_cutoff=args.intra_cutoff, normalize_coord=lambda x: x / scale, unnormalize_coord=lambda x: x * scale)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
        n = patch_reference_dropout(m, L, dseed, pdrop)
        m.train()
        b = batch_from_recipe(hidden, bkw, OUT)
        with torch.no_grad():
            X, H, pair = m(**b.clone().forward_args())
            m.eval()
            Xe, He, _ = m(**b.clone().forward_args())
        torch.save({"recipe": dict(hidden=hidden, n_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, far_ligand=False,
                                   flavour="plus", dropout_p=pdrop, dropout_seed=dseed, patched_dropouts=n),
                    "shapes": shapes, "X": X.clone(), "H": H.clone(), "pair": pair.clone(), "edges": [],
                    "torch": torch.__version__}, os.path.join(OUT, name + ".pt"))
        print(name, "patched", n, "dropouts; train-vs-eval deviation H", float((H - He).abs().max()), "X", float((X - Xe).abs().max()))


def main_l2_plus_sampling():
    """FABindPlus.inference in the reference's sampling mode: train() with the ranking modules in eval (test_sampling_fabind.py:
    118-124), every nn.Dropout patched to a column-only mask, DBSCAN clustering of the pocket centre and the confidence head on,
    python `random` seeded (cluster choice + random_n_iter draws)."""
    import random
    import torch.nn as nn
    from fabind_b200.dropout import keep_mask
    from fabind_b200.plus.model import HEAD_SITES
    from fabind_b200.synthetic import make_docking_batch
    mods = ref_shims.load_reference_model_module("plus")
    emb, pemb, L, IT, bkw, wseed, dseed, pdrop = 64, 32, 2, 2, dict(n_complexes=3, seed=1), 53, 4321, 0.1
    args = ref_shims.published_args_plus(mean_layers=L, n_iter=IT, dropout=pdrop, confidence_training=True, stack_mlp=True,
                                         use_clustering=True, random_n_iter=True)
    m = mods.model.FABindPlus(args, emb, pemb)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
    n = patch_reference_dropout(m, L, dseed, pdrop, prefix="complex_model")
    n += patch_reference_dropout(m, args.pocket_pred_layers, (dseed + 0x51ED27) & 0xFFFFFFFF, pdrop, prefix="pocket_pred_model")

    class HeadDrop(nn.Module):
        def __init__(self, site):
            super().__init__()
            self.site = site

        def forward(self, x):
            return x * keep_mask(dseed, self.site, 1, x.shape[-1], pdrop, colonly=True)[0] if self.training else x
    for name, site in HEAD_SITES.items():
        getattr(m, name).dropout = HeadDrop(site)
        n += 1
    m.train()
    for name, sub in m.named_modules():
        if name.startswith("confidence") or name.startswith("ranking"):
            sub.eval()
    d = make_docking_batch(**bkw)
    random.seed(99)
    with torch.no_grad():
        coords, batch, conf = m.inference(d.clone())
    torch.save({"recipe": dict(emb=emb, pemb=pemb, mean_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, dropout_p=pdrop,
                               dropout_seed=dseed, random_seed=99, patched_dropouts=n),
                "shapes": shapes, "coords": coords.clone(), "confidence": conf.clone(), "torch": torch.__version__},
               os.path.join(OUT, "l2plussample_h64_p32_l2_it2.pt"))
    print("l2plussample", "patched", n, "coords", tuple(coords.shape), "confidence", conf.tolist())


def main_post_optim():
    """post_optimize_compound_coords (utils/post_optim_utils.py:36-64) run unmodified on synthetic ligands"""
    import numpy as np
    from fabind_b200.synthetic import _one_complex
    ref = ref_shims.load_reference_post_optim()
    cases = []
    for n, shift, epochs, rigid in ((24, 0.0, 10, False), (24, 0.0, 10, True), (40, 0.0, 10, False), (24, 0.0, 100, False),
                                    (24, 0.0, 100, True), (18, 12.0, 1000, False), (40, 30.0, 1000, False), (60, 30.0, 1000, False)):
        rng = np.random.default_rng(100 + n + int(shift) + epochs + rigid)
        _, lig, _, las, lig_ref = _one_complex(rng, n, 30)
        refc = torch.tensor(lig_ref, dtype=torch.float32)
        pred = torch.tensor(lig + rng.normal(scale=0.8, size=lig.shape) + shift, dtype=torch.float32)
        las_t = torch.tensor(las.T.copy(), dtype=torch.long)
        x, loss, rmsd = ref.post_optimize_compound_coords(refc, pred, total_epoch=epochs, LAS_edge_index=None if rigid else las_t)
        cases.append(dict(n=n, epochs=epochs, rigid=rigid, ref=refc, pred=pred, las=las_t, x=x.clone(), loss=float(loss), rmsd=float(rmsd)))
        print("postopt", n, epochs, rigid, "loss", loss, "rmsd", rmsd)
    torch.save({"cases": cases, "torch": torch.__version__}, os.path.join(OUT, "postopt_cases.pt"))


def main_l2_plus_train_forward():
    """FABindPlus.forward(data, stage=2) in train() mode (what test_sampling_fabind.py's validate() calls): column-only dropout
    masks as above plus F.gumbel_softmax replaced by the same formula on INJECTED gumbel samples (model.py:136-137)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from fabind_b200.dropout import keep_mask
    from fabind_b200.plus.model import HEAD_SITES
    from fabind_b200.synthetic import make_docking_batch
    mods = ref_shims.load_reference_model_module("plus")
    emb, pemb, L, IT, bkw, wseed, dseed, pdrop = 64, 32, 2, 2, dict(n_complexes=3, seed=1), 54, 777, 0.1
    args = ref_shims.published_args_plus(mean_layers=L, n_iter=IT, dropout=pdrop, random_n_iter=False)
    m = mods.model.FABindPlus(args, emb, pemb)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
    n = patch_reference_dropout(m, L, dseed, pdrop, prefix="complex_model")
    n += patch_reference_dropout(m, args.pocket_pred_layers, (dseed + 0x51ED27) & 0xFFFFFFFF, pdrop, prefix="pocket_pred_model")

    class HeadDrop(nn.Module):
        def __init__(self, site):
            super().__init__()
            self.site = site

        def forward(self, x):
            return x * keep_mask(dseed, self.site, 1, x.shape[-1], pdrop, colonly=True)[0] if self.training else x
    for name, site in HEAD_SITES.items():
        getattr(m, name).dropout = HeadDrop(site)
    m.train()
    d = make_docking_batch(**bkw)
    pbw = d['protein_whole'].batch
    g = torch.Generator().manual_seed(5)
    noise = -torch.empty((pbw.shape[0], 2)).exponential_(generator=g).log()           # flat [n_res, 2]
    B = int(pbw.max()) + 1
    counts = torch.bincount(pbw, minlength=B)
    pos = torch.arange(pbw.shape[0]) - (torch.cumsum(counts, 0) - counts)[pbw]
    dense = torch.zeros((B, int(counts.max()), 2))
    dense[pbw, pos] = noise

    def gumbel_softmax(logits, tau=1, hard=False, eps=1e-10, dim=-1):                  # torch/nn/functional.py, noise injected
        y_soft = ((logits + dense) / tau).softmax(dim)
        if hard:
            index = y_soft.max(dim, keepdim=True)[1]
            y_hard = torch.zeros_like(logits).scatter_(dim, index, 1.0)
            return y_hard - y_soft.detach() + y_soft
        return y_soft
    orig = F.gumbel_softmax
    F.gumbel_softmax = gumbel_softmax
    try:
        with torch.no_grad():
            d2 = d.clone()
            fwd = m(d2, stage=2)
    finally:
        F.gumbel_softmax = orig
    torch.save({"recipe": dict(emb=emb, pemb=pemb, mean_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, dropout_p=pdrop,
                               dropout_seed=dseed),
                "shapes": shapes, "noise": noise, "forward": [t.clone() if torch.is_tensor(t) else t for t in fwd],
                "coords_after": d2.coords.clone(), "torch": torch.__version__}, os.path.join(OUT, "l2plustrainfwd_h64_p32_l2_it2.pt"))
    print("l2plustrainfwd", [tuple(t.shape) if torch.is_tensor(t) else t for t in fwd])


def main_grad():
    """Parameter gradients of the UNMODIFIED reference stacks (eval mode = no dropout, autograd on) for a fixed linear read-out
    of (X, H[, pair]): pins the oracles' autograd path (refine_coord: only the last iteration carries gradients), the checker
    of the training path (BASELINE config 5) that later rounds build."""
    for flavour in ("v1", "plus"):
        mods = ref_shims.load_reference(flavour)
        args = ref_shims.published_args_plus(dropout=0.0) if flavour == "plus" else ref_shims.published_args()
        hidden, L, IT, bkw, wseed = 32, 1, 2, dict(n_complexes=2, seed=6, n_c_range=(5, 9), n_p_range=(14, 22)), 91
        scale = args.coordinate_scale
        m = mods.att_model.EfficientMCAttModel(args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, n_iter=IT,
                                               inter_cutoff=args.inter_cutoff, intra_cutoff=args.intra_cutoff,
                                               normalize_coord=lambda x: x / scale, unnormalize_coord=lambda x: x * scale).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
        b = batch_from_recipe(hidden, bkw, OUT)
        g = torch.Generator().manual_seed(17)
        out = m(**b.clone().forward_args())
        rx, rh = torch.randn(out[0].shape, generator=g), torch.randn(out[1].shape, generator=g)
        loss = (out[0] * rx).sum() + (out[1] * rh).sum()
        if flavour == "plus":
            rp = torch.randn(out[2].shape, generator=g) * 0.1
            loss = loss + (out[2] * rp).sum()
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
        torch.save({"recipe": dict(hidden=hidden, n_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, far_ligand=False,
                                   flavour=flavour, readout_seed=17),
                    "shapes": shapes, "loss": float(loss), "grads": grads, "torch": torch.__version__},
                   os.path.join(OUT, f"grad_{flavour}_h32_l1_it2.pt"))
        print("grad", flavour, "loss", float(loss), "params with grad", sum(v is not None for v in grads.values()), "/", len(grads))


def patch_reference_dropout_v1(model, L, seed, p):
    """The v1 twin of patch_reference_dropout: every nn.Dropout of the unmodified FABind stack (models/egnn.py:38,158,362;
    models/cross_att.py:114) becomes a deterministic COLUMN-ONLY mask from the library's mask function.  MC_E_GCL uses ONE Dropout
    module twice per call (edge_mlp output egnn.py:82 -> site edge2, node_mlp output egnn.py:106 -> site node2); MCAttEGNN uses its
    own twice per call (after linear_in :398, before linear_out :461)."""
    import re
    import torch.nn as nn
    from fabind_b200.dropout import keep_mask, site_id, iter_seed
    state = dict(it=-1)

    class ColDrop(nn.Module):
        def __init__(self, sites):
            super().__init__()
            self.sites, self.calls = sites, 0

        def forward(self, x):
            site = self.sites[self.calls % len(self.sites)]
            self.calls += 1
            if not self.training:
                return x
            m = keep_mask(iter_seed(seed, state["it"]), site, 1, x.shape[-1], p, colonly=True)[0]
            return x * m

    def pre_hook(mod, inp):
        state["it"] += 1
    model.gnn.register_forward_pre_hook(pre_hook)
    n = 0
    for name, mod in list(model.named_modules()):
        if not isinstance(mod, nn.Dropout):
            continue
        parent = model
        parts = name.split(".")
        for q in parts[:-1]:
            parent = getattr(parent, q)
        if name == "gnn.dropout":
            new = ColDrop([site_id(-1, "stack_in"), site_id(-1, "stack_out")])
        else:
            m = re.match(r"gnn\.(gcl_(\d+)|att_(\d+)|out_layer)\.(.*)", name)
            assert m, name
            layer = L if m.group(1) == "out_layer" else int(m.group(2) or m.group(3))
            rest = m.group(4)
            if not m.group(1).startswith("att"):
                assert rest == "dropout", name
                new = ColDrop([site_id(layer, "edge2"), site_id(layer, "node2")])
            elif rest == "dropout":
                new = ColDrop([site_id(layer, "agg")])
            elif rest == "cross_attn_module.p_attention_block.dropout":
                new = ColDrop([site_id(layer, "patt")])
            elif rest == "cross_attn_module.c_attention_block.dropout":
                new = ColDrop([site_id(layer, "catt")])
            else:
                # dropout modules the published configuration constructs but never calls (e.g. inter_layer / pair blocks)
                new = ColDrop([0xFFFF])
        setattr(parent, parts[-1], new)
        n += 1
    return n


def main_grad_dropout():
    """Training-mode twin of main_grad for the v1 stack: the UNMODIFIED reference in train() mode (dropout 0.1 active in every
    refinement iteration, att_model.py:210-245), its nn.Dropout modules patched to the library's column-only masks; outputs, loss
    and parameter gradients for a fixed linear read-out.  Pins placement / scaling of every mask in the no_grad iterations, the
    training-mode forward AND the reverse pass of the training step (fabind_b200/train.py)."""
    mods = ref_shims.load_reference("v1")
    args = ref_shims.published_args()
    args.random_n_iter = False
    pdrop, dseed = 0.1, 4242
    for tag, hidden, L, IT, bkw, wseed in (("h32_l1_it2", 32, 1, 2, dict(n_complexes=2, seed=6, n_c_range=(5, 9), n_p_range=(14, 22)), 91),
                                           ("h64_l2_it3", 64, 2, 3, dict(n_complexes=3, seed=8, n_c_range=(6, 14), n_p_range=(20, 40)), 92)):
        scale = args.coordinate_scale
        m = mods.att_model.EfficientMCAttModel(args, hidden, hidden, 1, n_edge_feats=0, n_layers=L, dropout=pdrop, n_iter=IT,
                                               inter_cutoff=args.inter_cutoff, intra_cutoff=args.intra_cutoff,
                                               normalize_coord=lambda x: x / scale, unnormalize_coord=lambda x: x * scale)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(det_state_dict(shapes, wseed), strict=True)
        n = patch_reference_dropout_v1(m, L, dseed, pdrop)
        m.train()
        b = batch_from_recipe(hidden, bkw, OUT)
        g = torch.Generator().manual_seed(17)
        out = m(**b.clone().forward_args())
        rx, rh = torch.randn(out[0].shape, generator=g), torch.randn(out[1].shape, generator=g)
        loss = (out[0] * rx).sum() + (out[1] * rh).sum()
        loss.backward()
        grads = {k: (p.grad.clone() if p.grad is not None else None) for k, p in m.named_parameters()}
        torch.save({"recipe": dict(hidden=hidden, n_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, far_ligand=False, flavour="v1",
                                   readout_seed=17, dropout_p=pdrop, dropout_seed=dseed, dropout_colonly=True),
                    "shapes": shapes, "X": out[0].detach().clone(), "H": out[1].detach().clone(), "loss": float(loss), "grads": grads,
                    "torch": torch.__version__}, os.path.join(OUT, f"graddrop_v1_{tag}.pt"))
        print("graddrop v1", tag, "patched", n, "loss", float(loss), "params with grad", sum(v is not None for v in grads.values()), "/", len(grads))


def main_l2_plus():
    """goldens for the FABind+ L2 wrapper (FABind_plus/fabind/models/model.py::FABindPlus): forward(stage=2) in eval mode
    (13-tuple + the in-place shift of data.coords) and inference()"""
    from fabind_b200.synthetic import make_docking_batch
    mods = ref_shims.load_reference_model_module("plus")
    cases = {
        # published crop rule (radius = max(pred + 5, 20))
        "l2plus_h64_p32_l2_it2": (64, 32, 2, 2, dict(n_complexes=3, seed=1), 51, dict()),
        # per-complex predicted radius actually decides the crop (multiplicative buffer, no floor)
        "l2plus_h32_p32_l1_it2_radius": (32, 32, 1, 2, dict(n_complexes=2, seed=2, L_range=(120, 220)), 52,
                                         dict(pocket_radius_buffer=1.5, min_pocket_radius=0.0)),
    }
    for name, (emb, pemb, L, IT, bkw, wseed, over) in cases.items():
        args = ref_shims.published_args_plus(mean_layers=L, n_iter=IT, **over)
        m = mods.model.FABindPlus(args, emb, pemb).eval()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        sd = det_state_dict(shapes, wseed)
        if over:   # a radius head whose output lands in the 8-20 A range for these inputs
            sd["pocket_radius_head.linear2.bias"] = torch.full_like(sd["pocket_radius_head.linear2.bias"], 9.0)
        m.load_state_dict(sd, strict=True)
        d = make_docking_batch(**bkw)
        with torch.no_grad():
            d2 = d.clone()
            fwd = m(d2, stage=2)
            inf = m.inference(d.clone())
            d1 = d.clone()
            fwd1 = m(d1, stage=1)
        torch.save({"recipe": dict(emb=emb, pemb=pemb, mean_layers=L, n_iter=IT, batch=bkw, weight_seed=wseed, args_over=over,
                                   radius_bias=9.0 if over else None),
                    "shapes": shapes, "forward": [t.clone() if torch.is_tensor(t) else t for t in fwd],
                    "coords_after": d2.coords.clone(), "inference": inf[0].clone(),
                    "forward_stage1": [t.clone() if torch.is_tensor(t) else t for t in fwd1],
                    "coords_after_stage1": d1.coords.clone(), "complex_coords_after_stage1": d1['complex'].node_coords.clone(),
                    "torch": torch.__version__},
                   os.path.join(OUT, name + ".pt"))
        print(name, [tuple(t.shape) if torch.is_tensor(t) else t for t in fwd], "radius", fwd[11].flatten().tolist())


if __name__ == "__main__":
    which = sys.argv[1:] or ["v1", "l2", "plus", "l2plus", "plusdrop", "l2plussample", "postopt", "l2plustrainfwd", "grad", "graddrop"]
    if "l2plus" in which:
        main_l2_plus()
    if "plusdrop" in which:
        main_plus_dropout()
    if "l2plussample" in which:
        main_l2_plus_sampling()
    if "postopt" in which:
        main_post_optim()
    if "grad" in which:
        main_grad()
    if "graddrop" in which:
        main_grad_dropout()
    if "l2plustrainfwd" in which:
        main_l2_plus_train_forward()
    if "v1" in which:
        main()
    if "l2" in which:
        main_l2()
    if "plus" in which:
        main_plus()
