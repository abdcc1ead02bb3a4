"""Data-parallel sharding of a batch of complexes across ranks.

Complexes are independent units: every reduction on the docking path is within one complex (per-row segment ops, the
per-sample radial norm `egnn.py:775-779`, per-complex dense blocks), so the forward shards with NO data-path
collective.  Host logic only (runs with any torch.distributed backend; the CPU tests use gloo with world_size 2):

 * `partition(costs, world)`       size-balanced assignment (longest-processing-time first) of complexes to ranks;
 * `take_complexes(args, ids)`     the sub-batch holding the chosen complexes, node ids / edge lists re-based, in the
                                   argument layout of `EfficientMCAttModel.forward` (att_model.py:170);
 * `sharded_forward(model, args)`  each rank runs its shard through `model`, results are re-assembled in the caller's
                                   node order on every rank with one all_gather of (X, H) - outside the data path;
 * `max_over_ranks(ms)`            the timing reduction the benchmark contract asks for.
"""
import torch
import torch.distributed as dist


def complex_cost(n_c, n_p):
    """Work estimate of one complex: pair rows (Np' * Nc') + context edges (~10 per residue + bonds) + nodes."""
    return (n_p + 1) * (n_c + 1) + 12 * n_p + 4 * n_c


def partition(costs, world):
    """LPT greedy: returns `world` sorted lists of complex indices; every complex appears exactly once."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in out]


def _complex_ranges(batch_id):
    B = int(batch_id[-1]) + 1 if batch_id.numel() else 0
    counts = torch.bincount(batch_id, minlength=B)
    ends = torch.cumsum(counts, 0)
    return B, (ends - counts).tolist(), ends.tolist()


def take_complexes(fa, ids):
    """`fa`: dict with the reference forward's argument names (X, H, batch_id, segment_id, mask, is_global,
    compound_edge_index, LAS_edge_index, batched_complex_coord_LAS).  Returns (sub-batch dict, node index tensor)
    where node index maps sub-batch rows back to rows of the full batch."""
    bid = fa["batch_id"]
    B, starts, ends = _complex_ranges(bid)
    dev = bid.device
    node_idx = torch.cat([torch.arange(starts[i], ends[i], device=dev) for i in ids]) if ids else torch.zeros(0, dtype=torch.long, device=dev)
    new_of_old = torch.full((bid.numel(),), -1, dtype=torch.long, device=dev)
    new_of_old[node_idx] = torch.arange(node_idx.numel(), device=dev)
    new_bid = torch.full((B,), -1, dtype=torch.long, device=dev)
    new_bid[torch.tensor(ids, dtype=torch.long, device=dev)] = torch.arange(len(ids), device=dev)

    def edges(e):
        keep = new_of_old[e[0]] >= 0
        return new_of_old[e[:, keep]].contiguous()
    sub = dict(X=fa["X"][node_idx].contiguous(), H=fa["H"][node_idx].contiguous(), batch_id=new_bid[bid[node_idx]],
               segment_id=fa["segment_id"][node_idx], mask=fa["mask"][node_idx], is_global=fa["is_global"][node_idx],
               compound_edge_index=edges(fa["compound_edge_index"]), LAS_edge_index=edges(fa["LAS_edge_index"]),
               batched_complex_coord_LAS=fa["batched_complex_coord_LAS"][node_idx].contiguous(), LAS_mask=None)
    return sub, node_idx


def sharded_forward(model, fa, group=None):
    """Runs `model(**shard)` on this rank's complexes and returns (X, H) for the WHOLE batch on every rank.
    `model` follows the reference signature and returns (X, H[, ...]).  No collective touches the data path; the single
    all_gather at the end only re-assembles the outputs."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    bid = fa["batch_id"]
    B, starts, ends = _complex_ranges(bid)
    seg = fa["segment_id"].to(torch.bool)
    n_p = [int(seg[starts[i]:ends[i]].sum()) - 1 for i in range(B)]
    n_c = [ends[i] - starts[i] - n_p[i] - 2 for i in range(B)]
    parts = partition([complex_cost(c, p) for c, p in zip(n_c, n_p)], world)
    sub, node_idx = take_complexes(fa, parts[rank])
    if parts[rank]:
        out = model(**sub)
        Xs, Hs = out[0], out[1]
    else:
        Xs, Hs = sub["X"], fa["H"].new_zeros((0, fa["H"].shape[1]))
    N, hid = fa["X"].shape[0], Hs.shape[1] if Hs.numel() else fa["H"].shape[1]
    X_full = fa["X"].new_zeros(fa["X"].shape)
    H_full = fa["H"].new_zeros((N, hid))
    if world == 1:
        X_full[node_idx], H_full[node_idx] = Xs, Hs
        return X_full, H_full
    # ragged all_gather: pad every shard to the largest node count
    sizes = [sum(ends[i] - starts[i] for i in p) for p in parts]
    cap = max(sizes)
    pack = Xs.new_zeros((cap, 3 + hid))
    pack[:sizes[rank], :3] = Xs.reshape(-1, 3)
    pack[:sizes[rank], 3:] = Hs
    gathered = [torch.empty_like(pack) for _ in range(world)]
    dist.all_gather(gathered, pack, group=group)
    for r in range(world):
        idx = torch.cat([torch.arange(starts[i], ends[i], device=bid.device) for i in parts[r]]) if parts[r] else None
        if idx is None:
            continue
        X_full[idx] = gathered[r][:sizes[r], :3].reshape(-1, *fa["X"].shape[1:])
        H_full[idx] = gathered[r][:sizes[r], 3:]
    return X_full, H_full


def max_over_ranks(ms, device=None, group=None):
    """max over ranks of a device-measured duration (milliseconds)"""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(ms)
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


# ---- training side of the data parallelism: ONE all-reduce of all gradients per step -------------------
def _aliased_flat(params, model):
    """the flat gradient buffer of the last training step (train._backward_half) if every parameter's .grad still is a view into it,
    else None"""
    hint = getattr(model, "_fb_flat_grad", None) if model is not None else None
    if hint is None:
        return None
    flat, layout = hint
    named = dict(model.named_parameters())
    want = {id(p) for p in params}
    seen = 0
    for k, o, n, shape in layout:
        p = named.get(k)
        if p is None:
            continue
        if id(p) not in want or p.grad is None or p.grad.dtype != torch.float32 or not p.grad.is_contiguous() or \
                p.grad.data_ptr() != flat.data_ptr() + 4 * o or p.grad.numel() != n:
            return None
        seen += 1
    return flat if seen == len(want) else None


def allreduce_gradients(params, group=None, average=True, buffer=None, model=None):
    """Sum (or average) the gradients of `params` over the ranks with a SINGLE collective on one flat fp32 buffer (the reference
    trains under DDP with `find_unused_parameters=True`, FABind/fabind/main_fabind.py:198-200: 36-45 M fp32 gradients, the unused
    `att_i.inter_layer.*` parameters contribute zeros).  Parameters whose `.grad` is None on this rank are treated as zero and
    receive the reduced value if any rank had one.  Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
    Returns the flat buffer (re-usable through `buffer=` to avoid re-allocation)."""
    params = [p for p in params if p.requires_grad]
    # train.overlap_allreduce: the stack's gradients were already reduced inside the reverse pass of this step
    reduced = getattr(model, "_fb_reduced", None) if model is not None else None
    if reduced:
        object.__setattr__(model, "_fb_reduced", None)
        params = [p for p in params if id(p) not in reduced]
    if not params:
        return buffer
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    flat = _aliased_flat(params, model)
    if flat is not None:
        # the training step left all gradients in ONE flat buffer that the .grad tensors alias: reduce it in place, no copies
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat.div_(world)
        return flat
    n = sum(p.numel() for p in params)
    dev = params[0].device
    if buffer is None or buffer.numel() != n or buffer.device != dev:
        buffer = torch.empty(n, dtype=torch.float32, device=dev)
    o = 0
    for p in params:
        k = p.numel()
        if p.grad is None:
            buffer[o:o + k].zero_()
        else:
            buffer[o:o + k].copy_(p.grad.reshape(-1))
        o += k
    if world > 1:
        dist.all_reduce(buffer, op=dist.ReduceOp.SUM, group=group)
        if average:
            buffer.div_(world)
    o = 0
    for p in params:
        k = p.numel()
        g = buffer[o:o + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        o += k
    return buffer
