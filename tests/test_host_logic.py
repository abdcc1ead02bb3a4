"""Host-side logic that needs no GPU: node layout, weight packing (both layouts), the dropout mask function, batch replication
for batched sampling, argument validation of the host-only C-ABI entry points."""
import ctypes as C

import numpy as np
import pytest
import torch

from fabind_b200 import _lib
from fabind_b200.dropout import keep_mask, drop_hash, threshold, site_id, iter_seed
from fabind_b200.layout import build_layout
from fabind_b200.synthetic import make_batch, make_docking_batch
from fabind_b200.weights import pack_state_dict, slots
from oracle.det_weights import det_state_dict
from oracle import ref_shims


def test_layout_is_a_type_sorted_permutation():
    b = make_batch(n_complexes=4, seed=2, n_c_range=(3, 9), n_p_range=(5, 20), embed=8)
    lay = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    o, blob = lay.offs, lay.blob.numpy()
    N = lay.N
    perm, inv = blob[o["perm"]:o["perm"] + N], blob[o["inv"]:o["inv"] + N]
    assert sorted(perm.tolist()) == list(range(N)) and np.array_equal(inv[perm], np.arange(N))
    seg = b.segment_id.numpy().astype(bool)
    assert not seg[perm[:lay.Nc_tot]].any() and seg[perm[lay.Nc_tot:]].all()            # compound side first, then protein side
    assert np.all(np.diff(b.batch_id.numpy()[perm[:lay.Nc_tot]]) >= 0)                   # caller order kept inside a side
    c_off, p_off = blob[o["c_off"]:o["c_off"] + lay.B + 1], blob[o["p_off"]:o["p_off"] + lay.B + 1]
    pair_base = blob[o["pair_base"]:o["pair_base"] + lay.B + 1]
    assert c_off[0] == 0 and c_off[-1] == lay.Nc_tot == p_off[0] and p_off[-1] == N
    assert np.array_equal(np.diff(pair_base), np.diff(c_off) * np.diff(p_off)) and pair_base[-1] == lay.P_total
    assert lay.cap_int == 2 * int((np.array(b.n_c) * np.array(b.n_p)).sum())


def test_layout_rejects_bad_batches():
    b = make_batch(n_complexes=2, seed=1, n_c=3, n_p=5, embed=8)
    with pytest.raises(ValueError):
        build_layout(b.batch_id.flip(0), b.segment_id, b.is_global, b.mask, "cpu")       # unsorted batch ids
    with pytest.raises(ValueError):
        build_layout(b.batch_id, torch.zeros_like(b.segment_id), b.is_global, b.mask, "cpu")   # no protein side


@pytest.mark.parametrize("flavour", [0, 1])
def test_weight_packing_fills_every_slot(flavour):
    """every slot the library declares is produced by the packer with the declared shape, from the reference's key set"""
    from fabind_b200 import EfficientMCAttModel as V1
    from fabind_b200.plus import EfficientMCAttModel as Plus
    hidden, L = 64, 2
    args = ref_shims.published_args_plus() if flavour else ref_shims.published_args()
    m = (Plus if flavour else V1)(args, hidden, hidden, 1, n_layers=L, n_iter=1, normalize_coord=lambda x: x / 5.0,
                                  unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 3)
    arena = pack_state_dict(sd, hidden, L, flavour)
    assert arena.dtype == torch.float32 and torch.isfinite(arena).all()
    for name, r, c, off in slots(hidden, L, flavour):
        blk = arena[off:off + r * c]
        if not any(t in name for t in ("ca_c_b", "ca_p_b")):       # these carry structural zeros only in part
            assert float(blk.abs().sum()) > 0, name


def test_dropout_mask_function():
    assert threshold(0.1) == int(float(np.float32(0.1)) * 2 ** 32) and threshold(0.0) == 0
    m = keep_mask(11, site_id(2, "pair1"), 4000, 256, 0.1)
    assert abs(float((m > 0).float().mean()) - 0.9) < 0.005 and abs(float(m.max()) - 1 / 0.9) < 1e-6
    rows = (m > 0).float().mean(1)
    assert float(rows.std()) < 0.03                                              # no row / column structure
    assert not torch.equal(m, keep_mask(12, site_id(2, "pair1"), 4000, 256, 0.1))         # seed matters
    assert not torch.equal(m, keep_mask(11, site_id(2, "pair2"), 4000, 256, 0.1))         # site matters
    assert not torch.equal(m, keep_mask(iter_seed(11, 1), site_id(2, "pair1"), 4000, 256, 0.1))   # iteration matters
    assert torch.equal(keep_mask(11, 5, 10, 64, 0.3, row0=7)[0], keep_mask(11, 5, 20, 64, 0.3)[7])  # row0 = offset of a row block
    h = drop_hash(1, 2, np.arange(8), np.arange(8))
    assert h.dtype == np.uint64 and int(h.max()) < 2 ** 32


def test_replicate_batch_is_consistent():
    from fabind_b200.plus.sampling import replicate_batch
    d = make_docking_batch(2, seed=3, n_c_range=(4, 8), L_range=(30, 50))
    r = replicate_batch(d, 3)
    B = 2
    n_atoms, n_res, n_wp = d['compound'].batch.shape[0], d['protein_whole'].batch.shape[0], d['complex_whole_protein'].batch.shape[0]
    assert r['compound'].batch.tolist() == sum([[int(v) + k * B for v in d['compound'].batch] for k in range(3)], [])
    assert torch.equal(r['compound'].node_feats[n_atoms:2 * n_atoms], d['compound'].node_feats)
    assert torch.equal(r.node_xyz_whole[2 * n_res:], d.node_xyz_whole)
    e, e0 = r['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index, d['complex_whole_protein', 'c2c', 'complex_whole_protein'].edge_index
    assert torch.equal(e[:, e0.shape[1]:2 * e0.shape[1]], e0 + n_wp)
    wb = r['complex_whole_protein'].batch
    assert torch.equal(wb[e[0]], wb[e[1]])                                        # every bond stays inside its replica
    assert torch.equal(r['compound_atom_edge_list'].x[:d['compound_atom_edge_list'].x.shape[0]], d['compound_atom_edge_list'].x)


def test_host_entry_points_validate_arguments():
    l = _lib.lib()
    p = _lib.ModelParams()
    assert l.fb_graph_workspace_bytes(C.byref(p)) == -1 and l.fb_model_workspace_bytes(C.byref(p)) == -1      # all-zero params
    p.N, p.B, p.Nc_tot, p.P_total, p.hidden, p.n_layers, p.n_iter = 100, 2, 20, 800, 64, 2, 2
    p.cap_int, p.E_ctx = 3200, 900
    assert l.fb_graph_workspace_bytes(C.byref(p)) > 0 and l.fb_model_workspace_bytes(C.byref(p)) > 0
    plus = l.fb_model_workspace_bytes(C.byref(p))
    p.flavour = _lib.FLAVOUR_PLUS
    assert l.fb_model_workspace_bytes(C.byref(p)) > plus                          # FABind+ keeps pair ping-pong buffers
    p.hidden = 1024
    assert l.fb_model_workspace_bytes(C.byref(p)) == -1                           # hidden > 512 is not built
    p.hidden, p.flavour, p.dropout_p = 64, _lib.FLAVOUR_V1, 0.1
    assert l.fb_model_workspace_bytes(C.byref(p)) > 0                             # ABI 4: the v1 stack carries its training-mode dropout
    p.dropout_p = 1.0
    assert l.fb_model_workspace_bytes(C.byref(p)) == -1
    p.dropout_p, p.bf16_mode = 0.0, 7
    assert l.fb_model_workspace_bytes(C.byref(p)) == -1                           # unknown precision mode
    p.bf16_mode = _lib.PREC_SPLIT6
    fp32_tc = l.fb_model_workspace_bytes(C.byref(p))
    p.bf16_mode = _lib.PREC_FP32
    assert fp32_tc > l.fb_model_workspace_bytes(C.byref(p))                       # split modes add the operand scratch
    assert l.fb_weight_slot_info_f(64, 2, 1, 10 ** 6, None, 0, None, None, None) == -1
    assert l.fb_gemm(None, None) == -1 and l.fb_gemm_dot_tiles(45000, 512, 512, 1, 0) == 4


def test_apply_gradients_sets_param_grads_single_process():
    """train.apply_gradients: reference-shaped gradients land on the drop-in module's parameters (missing ones stay None -> zeros in
    the flat all-reduce buffer); world size 1: no collective"""
    import torch
    from fabind_b200 import EfficientMCAttModel, train
    from fabind_b200.config import published_args
    m = EfficientMCAttModel(published_args(), 32, 32, 1, n_layers=1, n_iter=1, normalize_coord=lambda x: x / 5.0,
                            unnormalize_coord=lambda x: x * 5.0)
    names = [k for k, _ in m.named_parameters()]
    pg = {k: torch.full_like(p, 0.5) for k, p in list(m.named_parameters())[:-2]}
    train.apply_gradients(m, pg)
    params = dict(m.named_parameters())
    assert all(float(params[k].grad.mean()) == 0.5 for k in names[:-2])
    assert all(params[k].grad is not None and float(params[k].grad.abs().max()) == 0.0 for k in names[-2:])


def test_fast_packer_matches_the_generic_packer():
    """weights.FastPackerV1 (what the training step uses every step: one gather + seven bilinear blocks per MC_Att_L, hand-written
    chain rule) against pack_state_dict / arena_grads_to_state_dict (the differentiable derivations): identical arena, same
    parameter gradients"""
    import torch
    from fabind_b200 import EfficientMCAttModel
    from fabind_b200.config import published_args
    from fabind_b200.weights import FastPackerV1, pack_state_dict, arena_grads_to_state_dict
    for H, L in [(64, 2), (128, 1)]:
        torch.manual_seed(H)
        m = EfficientMCAttModel(published_args(), H, H, 1, n_layers=L, n_iter=2, normalize_coord=lambda x: x / 5.0,
                                unnormalize_coord=lambda x: x * 5.0)
        sd = {k: v.detach() for k, v in m.state_dict().items()}
        for v in sd.values():
            v.copy_(torch.randn_like(v))
        fp = FastPackerV1(sd, H, L, "cpu")
        a0, a1 = pack_state_dict(sd, H, L, 0), fp.pack()
        assert torch.equal(a0, a1)
        ga = torch.randn_like(a0)
        g0 = arena_grads_to_state_dict(sd, ga, H, L, 0)
        flat, g1 = fp.unpack(ga)
        assert set(g0) == set(g1) and flat.numel() == sum(v.numel() for v in g0.values())
        for k in g0:
            assert float((g0[k] - g1[k]).abs().max()) <= 2e-6 * float(g0[k].abs().max()) + 1e-7, k
        # the packer rewrites ONE arena / parameter buffer in place: a second pack after an optimizer step sees the new weights
        for v in sd.values():
            v.mul_(0.5).add_(0.01)
        a2 = fp.pack()
        assert a2.data_ptr() == a1.data_ptr() and torch.equal(a2, pack_state_dict(sd, H, L, 0))
        flat2, g2 = fp.unpack(ga)
        g0b = arena_grads_to_state_dict(sd, ga, H, L, 0)
        for k in g0b:
            assert float((g0b[k] - g2[k]).abs().max()) <= 2e-6 * float(g0b[k].abs().max()) + 1e-7, k


def test_layout_cache_and_moving_rows_count():
    """build_layout is cached per batch object (same tensors, unchanged versions) and counts the masked rows (fb_model_params.n_mv)"""
    from fabind_b200.synthetic import make_batch
    from fabind_b200.layout import build_layout
    b = make_batch(n_complexes=3, n_c=12, n_p=40, embed=32, seed=4)
    l1 = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    assert build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu") is l1
    assert l1.n_mv == int(b.mask.sum()) == 3 * (12 + 2)          # ligand atoms + both global nodes (model.py:261-262)
    b.mask[0] = ~b.mask[0]                                        # in-place edit: version bump -> rebuilt
    l2 = build_layout(b.batch_id, b.segment_id, b.is_global, b.mask, "cpu")
    assert l2 is not l1 and l2.n_mv == int(b.mask.sum())
    l3 = build_layout(b.batch_id.clone(), b.segment_id, b.is_global, b.mask, "cpu")     # another tensor object: rebuilt
    assert l3 is not l2 and l3.n_mv == l2.n_mv


def test_derived_slots_are_hidden_from_the_packers():
    """weights.slots() lists the library's derived (f_*) slots only on request; the packed arena leaves them zero (fb_derive_weights
    fills them on the device) and all slots stay disjoint"""
    import torch
    from fabind_b200 import _lib
    from fabind_b200.weights import slots
    H, L = 64, 2
    base, full = slots(H, L, 0), slots(H, L, 0, derived=True)
    names = {n for n, *_ in base}
    extra = [(n, r, c, o) for n, r, c, o in full if n not in names]
    assert extra and all(n.rpartition(".")[2].startswith("f_") for n, *_ in extra)
    assert not any(n.rpartition(".")[2].startswith("f_") for n in names)
    end = 0
    for n, r, c, o in full:
        assert o >= end, n
        end = o + r * c
    assert end <= _lib.lib().fb_weight_arena_elems_f(H, L, 0)
    from fabind_b200.weights import base_elems
    assert all(o >= base_elems(H, L, 0) for _, _, _, o in extra)  # derived slots sit behind the base prefix of the arena
    assert slots(H, L, 1, derived=True) == slots(H, L, 1)        # FABind+ layout: no folded slots
