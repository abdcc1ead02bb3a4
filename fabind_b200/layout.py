"""Host-side batch layout: the internal, type-sorted node order of libfabind_b200 (see csrc/graph.h).

All compound-side nodes (segment 0: glb_c + ligand atoms) of all complexes come first, then all
protein-side nodes (segment 1: glb_p + residues); inside a side nodes keep the caller's order, so the
dense per-complex blocks the reference builds with to_dense_batch (egnn.py:260-265) are row slices.
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Layout:
    N: int
    B: int
    Nc_tot: int
    P_total: int
    cap_int: int
    fb_atom: int
    fb_res: int
    max_c: int
    max_p: int
    blob: torch.Tensor      # int32 device blob holding perm|inv|node_cplx|c_off|p_off|pair_base
    flags: torch.Tensor     # uint8 device [N]
    offs: dict              # name -> element offset in blob
    orig_off: np.ndarray    # [B+1] node offsets in caller order
    n_c: np.ndarray
    n_p: np.ndarray
    n_mv: int = 0                   # number of masked (moving) nodes
    # dataloader-side edge counts (fabind_b200/dataloader.py): context edges incl. `hint_n_bond` bonds, and those INTO the moving
    # rows; None = unknown (the runtime reads them back after fb_graph_static: one host sync per forward)
    e_ctx: int = None
    e_ctx_mv: int = None
    hint_n_bond: int = None
    hint_cutoff: float = None

    def ptr(self, name):
        return self.blob.data_ptr() + 4 * self.offs[name]


_CACHE = []          # [(weakrefs of the four index tensors, their versions, device, allow_single_side, Layout)], most recent last
_CACHE_MAX = 8


def _norm_device(device):
    dev = torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def build_layout(batch_id, segment_id, is_global, mask, device, allow_single_side=False):
    """Layout of a batch, cached per batch OBJECT: the sampling protocol (40 passes over the same `data`, P/test_sampling_fabind.py:
    118-131), evaluation loops and benches hand the same four index tensors to every forward, and rebuilding the layout costs a
    device->host transfer (one sync) plus the numpy pass each time.  A hit requires the very same tensor objects (weak references,
    still alive) with unchanged `_version`s on the same target device; in-place edits bump the version, new tensors miss.  (Writes
    through `.data` are invisible to this test, as they are to autograd.)"""
    import weakref
    dev = _norm_device(device)
    tensors = (batch_id, segment_id, is_global, mask)
    vers = tuple(t._version for t in tensors)
    for i in range(len(_CACHE) - 1, -1, -1):
        refs, v, d, a, lay = _CACHE[i]
        if d == dev and a == allow_single_side and v == vers and all(r() is t for r, t in zip(refs, tensors)):
            return lay
    lay = _build_layout(batch_id, segment_id, is_global, mask, dev, allow_single_side)
    try:
        _CACHE.append((tuple(weakref.ref(t) for t in tensors), vers, dev, allow_single_side, lay))
        del _CACHE[:-_CACHE_MAX]
    except TypeError:
        pass
    return lay


def _build_layout(batch_id, segment_id, is_global, mask, device, allow_single_side=False):
    return materialize(layout_arrays(batch_id, segment_id, is_global, mask, allow_single_side), device)


def register(tensors, lay, device, allow_single_side=False):
    """Enter a layout made elsewhere (fabind_b200.dataloader: in a collate function / loader worker, with the edge counts of the
    context graph) for the four index tensor OBJECTS (batch_id, segment_id, is_global, mask) a later forward will be handed."""
    import weakref
    tensors = tuple(tensors)
    _CACHE.append((tuple(weakref.ref(t) for t in tensors), tuple(t._version for t in tensors), _norm_device(device), allow_single_side, lay))
    del _CACHE[:-_CACHE_MAX]
    return lay


def layout_arrays(batch_id, segment_id, is_global, mask, allow_single_side=False):
    """The host half of the layout: numpy only (safe in a DataLoader worker, picklable result)."""
    if batch_id.device.type != "cpu" and all(t.device == batch_id.device for t in (segment_id, is_global, mask)):
        # device inputs: ONE device->host transfer (one sync) for the four index vectors instead of four
        packed = (batch_id.detach().to(torch.int64) * 8 + segment_id.detach().to(torch.int64) + is_global.detach().to(torch.int64) * 2
                  + mask.detach().to(torch.int64) * 4).cpu().numpy()
        bid, seg, glb, msk = packed >> 3, (packed & 1).astype(bool), (packed & 2).astype(bool), (packed & 4).astype(bool)
    else:
        bid = batch_id.detach().cpu().numpy().astype(np.int64)
        seg = segment_id.detach().cpu().numpy().astype(bool)
        glb = is_global.detach().cpu().numpy().astype(bool)
        msk = mask.detach().cpu().numpy().astype(bool)
    N = bid.shape[0]
    if N == 0:
        raise ValueError("empty batch")
    if np.any(np.diff(bid) < 0):
        raise ValueError("batch_id must be sorted (as torch_geometric batches are)")
    B = int(bid[-1]) + 1
    counts = np.bincount(bid, minlength=B)
    if np.any(counts == 0):
        raise ValueError("every complex id in [0, B) must own at least one node")
    orig_off = np.concatenate([[0], np.cumsum(counts)])
    c_idx = np.nonzero(~seg)[0]
    p_idx = np.nonzero(seg)[0]
    perm = np.concatenate([c_idx, p_idx]).astype(np.int32)
    inv = np.empty(N, dtype=np.int32)
    inv[perm] = np.arange(N, dtype=np.int32)
    nc1 = np.bincount(bid[c_idx], minlength=B)
    np1 = np.bincount(bid[p_idx], minlength=B)
    if (np.any(nc1 == 0) or np.any(np1 == 0)) and not allow_single_side:
        raise ValueError("every complex needs compound-side and protein-side nodes")
    if int((nc1.astype(np.int64) * np1.astype(np.int64)).sum()) * 2 >= 2 ** 31:
        raise ValueError("batch too large for the library's 32-bit row indices: split it (fabind_b200.shard.take_complexes)")
    Nc_tot = int(nc1.sum())
    c_off = np.concatenate([[0], np.cumsum(nc1)]).astype(np.int32)
    p_off = (Nc_tot + np.concatenate([[0], np.cumsum(np1)])).astype(np.int32)
    pair_base = np.concatenate([[0], np.cumsum(nc1 * np1)]).astype(np.int32)
    node_cplx = bid[perm].astype(np.int32)
    flags = (seg[perm].astype(np.uint8) | (glb[perm].astype(np.uint8) << 1) | (msk[perm].astype(np.uint8) << 2))
    n_c = np.bincount(bid[(~seg) & (~glb)], minlength=B)
    n_p = np.bincount(bid[seg & (~glb)], minlength=B)
    cap_int = max(2, int(2 * (n_c * n_p).sum()))
    # zero-inter-edge fallback pair (att_model.py:85-86): first non-global compound / protein node of complex 0
    c0 = [i for i in range(c_off[0], c_off[1]) if not (flags[i] & 2)]
    p0 = [i for i in range(p_off[0], p_off[1]) if not (flags[i] & 2)]
    fb_atom = c0[0] if c0 else -1
    fb_res = p0[0] if p0 else -1
    parts = dict(perm=perm, inv=inv, node_cplx=node_cplx, c_off=c_off, p_off=p_off, pair_base=pair_base)
    offs, cur, chunks = {}, 0, []
    for k, v in parts.items():
        offs[k] = cur
        pad = (-len(v)) % 4
        chunks.append(np.concatenate([v, np.zeros(pad, np.int32)]))
        cur += len(v) + pad
    return dict(N=N, B=B, Nc_tot=Nc_tot, P_total=int(pair_base[-1]), cap_int=cap_int, fb_atom=int(fb_atom), fb_res=int(fb_res),
                max_c=int(nc1.max()), max_p=int(np1.max()) if len(np1) else 0, blob=np.concatenate(chunks), flags=flags, offs=offs,
                orig_off=orig_off, n_c=n_c, n_p=n_p, n_mv=int(msk.sum()))


def materialize(arr, device, e_ctx=None, e_ctx_mv=None, hint_n_bond=None, hint_cutoff=None):
    """numpy layout (layout_arrays) -> Layout with its two index blobs on `device` (asynchronous copies, no sync)"""
    arr = dict(arr)
    blob, flags = torch.from_numpy(arr.pop("blob")), torch.from_numpy(arr.pop("flags"))
    if torch.device(device).type == "cuda":
        blob, flags = blob.pin_memory(), flags.pin_memory()
    return Layout(blob=blob.to(device, non_blocking=True), flags=flags.to(device, non_blocking=True), e_ctx=e_ctx, e_ctx_mv=e_ctx_mv,
                  hint_n_bond=hint_n_bond, hint_cutoff=hint_cutoff, **arr)
