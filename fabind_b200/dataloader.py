"""Dataloader-side batch layout (reference FABind/fabind/utils/utils.py:202-442 assembles each complex on
the CPU, models/att_model.py:38-116 rebuilds the context graph on the device every iteration).

What a collate function (or a DataLoader worker) can know without touching the GPU:
  * the type-sorted node order, per-complex offsets and node flags of the library (`layout.layout_arrays`, numpy only);
  * the SIZE of the context graph.  Context edges are the ligand bonds, the residue-residue pairs within `intra_cutoff` and the
    global-node edges (att_model.py:64-107); protein rows and the bond list never change between refinement iterations
    (att_model.py:232-236 moves the masked nodes only), so the count depends on the batch alone.  With it the host sizes the
    edge-level scratch and the GEMM row counts, and the forward needs NO device->host read between its entry and its outputs
    (fb_model_params.layout_flag; the device still builds the CSR lists itself and checks the claim).

`layout_hint(...)` is the worker half (CPU tensors in, picklable `LayoutHint` out); `attach(hint, tensors, device)` is the
main-process half: it uploads the two small index blobs (pinned, asynchronous) and enters the layout for the tensor OBJECTS the
forward will be handed.  `prepare_batch` does both for a dict of forward arguments and moves it to the device.

Distance predicate: the reference's own expression on the CPU, `torch.norm(x_i - x_j, dim=-1) <= cutoff` (att_model.py:124-126) --
the device predicate of csrc/graph.cu is bit-identical to it (tests/test_gpu_edges.py), so the counts agree even for pairs that sit
on the cutoff."""
from dataclasses import dataclass

import numpy as np
import torch

from . import layout as _layout


@dataclass
class LayoutHint:
    arrays: dict        # layout.layout_arrays(...)
    e_ctx: int          # context edges: bonds + residue-residue within the cutoff + global-node edges
    e_ctx_mv: int       # context edges whose row (receiving node) is a masked node
    n_bond: int
    intra_cutoff: float


def _pp_degrees(xp, cutoff):
    """number of OTHER rows of xp [n,3] within `cutoff` of each row (the reference's CPU predicate, att_model.py:124-126)"""
    n = xp.shape[0]
    if n < 2:
        return np.zeros(n, dtype=np.int64)
    deg = np.zeros(n, dtype=np.int64)
    step = max(1, (1 << 22) // n)                      # <= 4M pairs per block
    for i0 in range(0, n, step):
        a = xp[i0:i0 + step]
        d = (a[:, None, :] - xp[None, :, :]).reshape(-1, 3).contiguous()
        hit = (torch.norm(d, dim=-1) <= cutoff).view(a.shape[0], n)
        idx = torch.arange(i0, i0 + a.shape[0])
        hit[torch.arange(a.shape[0]), idx] = False     # no self loops (att_model.py:59)
        deg[i0:i0 + a.shape[0]] = hit.sum(1).numpy()
    return deg


def context_degrees(X, batch_id, segment_id, is_global, bonds, intra_cutoff):
    """Context-graph degree of every node (caller order), i.e. the number of context edges whose ROW is that node: the categories
    of csrc/graph.cu::edge_category restated on the host -- bond edges by their first endpoint; residue-residue pairs within the
    cutoff; every ordered pair of one segment of one complex with at least one global node; the global-global pairs across the
    two segments (att_model.py:64-107)."""
    x = X.detach().reshape(-1, 3).to(torch.float32).cpu()
    bid = batch_id.detach().cpu().numpy().astype(np.int64)
    seg = segment_id.detach().cpu().numpy().astype(bool)
    glb = is_global.detach().cpu().numpy().astype(bool)
    N = bid.shape[0]
    B = int(bid[-1]) + 1
    deg = np.zeros(N, dtype=np.int64)
    if bonds is not None and bonds.numel():
        deg += np.bincount(bonds[0].detach().cpu().numpy().astype(np.int64), minlength=N)
    # per (complex, segment): total nodes and global nodes
    key = bid * 2 + seg
    n_tot = np.bincount(key, minlength=2 * B)
    n_glb = np.bincount(key[glb], minlength=2 * B)
    g_other = n_glb[bid * 2 + (1 - seg.astype(np.int64))]
    deg += np.where(glb, n_tot[key] - 1 + g_other, n_glb[key])
    off = np.concatenate([[0], np.cumsum(np.bincount(bid, minlength=B))])
    cutoff = float(intra_cutoff)
    for b in range(B):
        idx = off[b] + np.nonzero(seg[off[b]:off[b + 1]] & ~glb[off[b]:off[b + 1]])[0]
        if len(idx) > 1:
            deg[idx] += _pp_degrees(x[torch.from_numpy(idx)], cutoff)
    return deg


def layout_hint(X, batch_id, segment_id, mask, is_global, compound_edge_index, intra_cutoff, allow_single_side=False):
    """CPU tensors of one collated batch (the forward's own arguments; X in the model's NORMALISED units, `intra_cutoff` =
    normalize_coord(8), i.e. `model.layout_cutoff()`) -> LayoutHint.  numpy / torch-CPU only."""
    arrays = _layout.layout_arrays(batch_id, segment_id, is_global, mask, allow_single_side)
    deg = context_degrees(X, batch_id, segment_id, is_global, compound_edge_index, intra_cutoff)
    msk = mask.detach().cpu().numpy().astype(bool)
    n_bond = 0 if compound_edge_index is None else int(compound_edge_index.shape[1])
    return LayoutHint(arrays=arrays, e_ctx=int(deg.sum()), e_ctx_mv=int(deg[msk].sum()), n_bond=n_bond, intra_cutoff=float(intra_cutoff))


def attach(hint, batch_id, segment_id, is_global, mask, device, allow_single_side=False):
    """Main-process half: the layout of `hint` on `device`, entered for these four tensor objects (the ones the forward receives)."""
    lay = _layout.materialize(hint.arrays, device, e_ctx=hint.e_ctx, e_ctx_mv=hint.e_ctx_mv, hint_n_bond=hint.n_bond,
                              hint_cutoff=hint.intra_cutoff)
    return _layout.register((batch_id, segment_id, is_global, mask), lay, device, allow_single_side)


def prepare_batch(args, device, intra_cutoff, move=True, hint=None):
    """`args`: dict of CPU tensors with the forward's argument names (X, H, batch_id, segment_id, mask, is_global,
    compound_edge_index, LAS_edge_index, batched_complex_coord_LAS[, LAS_mask]); `device`: the model's device.  Returns the dict
    the forward is called with -- on `device` (pinned, asynchronous copies; `move=False` keeps the tensors on the host for the
    model's host-buffer mode) -- with its layout and context-edge counts attached: `model(**prepare_batch(...))` then runs without
    a device->host read between entry and outputs."""
    if hint is None:
        hint = layout_hint(args["X"], args["batch_id"], args["segment_id"], args["mask"], args["is_global"],
                           args["compound_edge_index"], intra_cutoff)
    dev = torch.device(device)
    out = dict(args)
    if move:
        out = {k: (v.pin_memory().to(dev, non_blocking=True) if torch.is_tensor(v) and v.device.type == "cpu" else v) for k, v in args.items()}
    attach(hint, out["batch_id"], out["segment_id"], out["is_global"], out["mask"], dev)
    return out
