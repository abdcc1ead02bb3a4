"""Host logic of the N>1 path (fabind_b200/shard.py) on CPU with the gloo backend, world_size 2: the partition is a
disjoint cover, re-based sub-batches are self-consistent, and the sharded forward re-assembles per-complex results in
the caller's order.  The per-rank `model` here is the CPU ORACLE acting as the checker's stand-in for the CUDA module
(the product path has no CPU mode); the GPU tests run the same function with the real module."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fabind_b200 import shard
from fabind_b200.synthetic import make_batch
from oracle import fabind_oracle as orc
from oracle.det_weights import det_state_dict
from helpers import golden_files

HID, L, IT = 32, 1, 2


def test_partition_is_balanced_disjoint_cover():
    costs = [shard.complex_cost(c, p) for c, p in [(30, 200), (10, 80), (80, 250), (45, 120), (12, 90), (60, 240), (33, 150)]]
    for world in (1, 2, 3, 8):
        parts = shard.partition(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)          # LPT bound


def test_take_complexes_rebases_ids():
    b = make_batch(n_complexes=4, seed=3, n_c_range=(5, 12), n_p_range=(20, 40), embed=HID)
    fa = b.forward_args()
    sub, idx = shard.take_complexes(fa, [1, 3])
    assert sub["batch_id"].unique().tolist() == [0, 1]
    assert torch.equal(sub["X"], fa["X"][idx]) and torch.equal(sub["H"], fa["H"][idx])
    # every kept bond joins two nodes of the same complex, inside range, and none was lost
    e = sub["compound_edge_index"]
    assert e.min() >= 0 and e.max() < idx.numel()
    assert torch.equal(sub["batch_id"][e[0]], sub["batch_id"][e[1]])
    kept = ((fa["batch_id"][fa["compound_edge_index"][0]] == 1) | (fa["batch_id"][fa["compound_edge_index"][0]] == 3)).sum()
    assert e.shape[1] == int(kept)
    assert torch.equal(idx[e], fa["compound_edge_index"][:, torch.isin(fa["batch_id"][fa["compound_edge_index"][0]], torch.tensor([1, 3]))])


def _oracle_model(sd, cfg):
    def model(X, H, batch_id, segment_id, mask, is_global, compound_edge_index, LAS_edge_index, batched_complex_coord_LAS,
              LAS_mask=None):
        with torch.no_grad():
            return orc.model_forward(sd, cfg, X, H, batch_id, segment_id, mask, is_global, compound_edge_index,
                                     LAS_edge_index, batched_complex_coord_LAS)
    return model


def _weights():
    import torch as t
    g = t.load(golden_files()[-1], map_location="cpu", weights_only=False)   # any v1 fixture with hidden 32: key/shape table
    for p in golden_files():
        g = t.load(p, map_location="cpu", weights_only=False)
        if g["recipe"]["hidden"] == HID and g["recipe"]["n_layers"] == L:
            return det_state_dict(g["shapes"], 77)
    raise RuntimeError("no hidden-32 fixture")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b = make_batch(n_complexes=5, seed=9, n_c_range=(5, 14), n_p_range=(20, 45), embed=HID)
        model = _oracle_model(_weights(), orc.make_cfg(n_layers=L, n_iter=IT))
        X, H = shard.sharded_forward(model, b.forward_args())
        t = shard.max_over_ranks(10.0 + rank)
        q.put((rank, X, H, t))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_forward_world2_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    b = make_batch(n_complexes=5, seed=9, n_c_range=(5, 14), n_p_range=(20, 45), embed=HID)
    Xr, Hr = _oracle_model(_weights(), orc.make_cfg(n_layers=L, n_iter=IT))(**b.forward_args())
    for rank, X, H, t in res:
        # complexes are independent: sharding must not change any complex's result beyond fp32 reduction order of the
        # per-sample norms (identical here: same code on the same rows)
        assert torch.allclose(X, Xr, atol=1e-6) and torch.allclose(H, Hr, atol=1e-5), rank
        assert t == 11.0            # max over ranks of (10, 11)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.zeros(7, 5)), torch.nn.Parameter(torch.zeros(11)), torch.nn.Parameter(torch.zeros(3, 3)),
              torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
        ps[0].grad = torch.full((7, 5), float(rank + 1))
        ps[1].grad = torch.arange(11.0) * (rank + 1)
        if rank == 1:
            ps[2].grad = torch.ones(3, 3)            # unused on rank 0 (find_unused_parameters semantics)
        buf = shard.allreduce_gradients(ps, average=True)
        q.put((rank, [p.grad.clone() if p.grad is not None else None for p in ps], buf.numel()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_gradient_allreduce_world2():
    """one flat all-reduce for all gradients: mean over ranks, missing gradients count as zero, frozen parameters untouched"""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, grads, n in res:
        assert n == 35 + 11 + 9
        assert torch.equal(grads[0], torch.full((7, 5), 1.5)) and torch.equal(grads[1], torch.arange(11.0) * 1.5)
        assert torch.equal(grads[2], torch.full((3, 3), 0.5)) and grads[3] is None


def _train_case(seed):
    """one small batch + a v1 module with deterministic weights + fixed output gradients"""
    from fabind_b200 import EfficientMCAttModel
    from fabind_b200.config import published_args
    b = make_batch(n_complexes=2, seed=seed, n_c_range=(5, 10), n_p_range=(14, 24), embed=HID)
    m = EfficientMCAttModel(published_args(), HID, HID, 1, n_layers=L, n_iter=IT, normalize_coord=lambda x: x / 5.0,
                            unnormalize_coord=lambda x: x * 5.0)
    m.load_state_dict(_weights(), strict=True)
    g = torch.Generator().manual_seed(seed + 100)
    return b, m, torch.randn(b.X.shape, generator=g), torch.randn(b.H.shape, generator=g)


def _rank_gradients(seed, overlap=False):
    """parameter gradients of one rank's batch through train.training_step (kernel wrappers -> torch stand-ins, providers -> oracle)"""
    from _pytest.monkeypatch import MonkeyPatch
    from fabind_b200 import backward as bw, train
    import test_backward_orchestration as T
    mp_ = MonkeyPatch()
    T._install_standins(mp_, bw)
    T._install_forward_standins(mp_, bw)
    mp_.setattr(bw, "pair_bias_gate_bwd", T._gate_bwd_standin)
    mp_.setattr(bw, "pair_outer_bwd", T._outer_bwd_standin)
    b, m, rx, rh = _train_case(seed)
    if overlap:
        train.overlap_allreduce(m, average=True)
    sd = _weights()
    cfg = orc.make_cfg(n_layers=L, n_iter=IT)

    def prev_coords(model, fa):
        with torch.no_grad():
            return orc.model_forward(sd, orc.make_cfg(n_layers=L, n_iter=IT - 1), fa["X"], fa["H"], fa["batch_id"], fa["segment_id"], fa["mask"],
                                     fa["is_global"], fa["compound_edge_index"], fa["LAS_edge_index"], fa["batched_complex_coord_LAS"])[0]

    def edge_lists(model, X_prev, fa):
        ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                        cfg.inter_cutoff / cfg.coordinate_scale)
        return ctx, inter
    try:
        _, _, pg, _ = train.training_step(m, b.forward_args(), lambda X, H: (rx, rh), prev_coords=prev_coords, edge_lists=edge_lists)
    finally:
        mp_.undo()
    return m, pg


def _train_worker(rank, world, port, q, overlap=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fabind_b200 import train
        if overlap:
            # the collective runs INSIDE the reverse pass, group by group (train.overlap_allreduce); apply_gradients' own all-reduce
            # must then leave the stack's parameters alone (they are already averaged)
            calls = []
            real = dist.all_reduce
            def counting(t, *a, **k):
                calls.append(t.numel())
                return real(t, *a, **k)
            dist.all_reduce = counting
            m, pg = _rank_gradients(20 + rank, overlap=True)
            n_inside = len(calls)
            assert n_inside >= 2 * L + 2, calls            # out layer, att_l / gcl_l per layer, top-level slots
            train.apply_gradients(m, pg, average=True)
            assert len(calls) == n_inside, "second reduction of already reduced gradients"
            dist.all_reduce = real
        else:
            m, pg = _rank_gradients(20 + rank)
            train.apply_gradients(m, pg, average=True)
        q.put((rank, {k: p.grad.numpy().copy() for k, p in m.named_parameters()}))   # numpy: no shared-memory handles that die with the worker
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("overlap", [False, True], ids=["tail_allreduce", "overlapped_allreduce"])
def test_data_parallel_training_step_world2(overlap):
    """config 5's structure on CPU: every rank differentiates ITS complexes (train.training_step), one flat all-reduce averages the
    gradients (train.apply_gradients); every rank ends with the mean of the per-rank gradients, unused parameters as zeros"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, overlap)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = [_rank_gradients(20 + r)[1] for r in range(2)]
    gmax = max(float((0.5 * (ref[0][k] + ref[1][k])).abs().max()) for k in ref[0])
    for rank, grads in res:
        for k, g in grads.items():
            want = 0.5 * (ref[0][k] + ref[1][k])
            # (the workers run with 2 threads, the parent with all: summation order differs in the last bits)
            assert float((torch.from_numpy(g) - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-6 * gmax, (rank, k)
    assert any(float(abs(g).max()) == 0.0 for g in res[0][1].values())          # att_i.inter_layer.*: never used, zero gradient
