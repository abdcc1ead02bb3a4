"""Drop-in for FABind/fabind/utils/post_optim_utils.py::post_optimize_compound_coords (and the per-ligand loop around it,
fabind_inference.py:285-316): the 1000-step Adam refinement of the predicted ligand coordinates against the LAS distance
constraints, for a whole BATCH of ligands in one kernel launch (`fb_post_optimize`, csrc/postopt.cu)."""
import ctypes as C

import torch

from . import _lib
from .runtime import current_stream_ptr


def post_optimize_batch(reference_coords, predict_coords, compound_batch, las_local=None, las_batch=None, total_epoch=1000, lr=0.1):
    """reference_coords / predict_coords: [n_atoms, 3] (all ligands concatenated, CUDA); compound_batch: [n_atoms] sorted ligand id;
    las_local: [2, E] ligand-LOCAL atom ids with las_batch [E] sorted ligand ids (None = rigid mode, post_optim_utils.py:31).
    Returns (coords [n_atoms, 3], loss [B], rmsd [B]) -- per ligand what the reference function returns."""
    l = _lib.lib()
    dev = predict_coords.device
    if dev.type != "cuda":
        raise RuntimeError("fabind_b200 runs on a CUDA device only (no CPU fallback)")
    ref = reference_coords.to(dev, torch.float32).contiguous()
    pred = predict_coords.to(torch.float32).contiguous()
    B = int(compound_batch.max()) + 1 if compound_batch.numel() else 0
    counts = torch.bincount(compound_batch.to(dev), minlength=B)
    off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    off[1:] = torch.cumsum(counts, 0).to(torch.int32)
    max_atoms = int(counts.max()) if B else 0
    out = torch.empty_like(pred)
    loss = torch.zeros(B, dtype=torch.float32, device=dev)
    rmsd = torch.zeros(B, dtype=torch.float32, device=dev)
    las_ptr = las_off_ptr = None
    n_las = 0
    if las_local is not None:
        las_d = las_local.to(dev, torch.int32).contiguous()
        n_las = las_d.shape[1]
        lc = torch.bincount(las_batch.to(dev), minlength=B)
        las_off = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        las_off[1:] = torch.cumsum(lc, 0).to(torch.int32)
        las_ptr, las_off_ptr = las_d.data_ptr(), las_off.data_ptr()
    _lib.check(l.fb_post_optimize(ref.data_ptr(), pred.data_ptr(), off.data_ptr(), B, max_atoms, las_ptr, las_off_ptr, n_las,
                                  int(total_epoch), float(lr), out.data_ptr(), loss.data_ptr(), rmsd.data_ptr(),
                                  current_stream_ptr(dev)), "fb_post_optimize")
    return out, loss, rmsd


def post_optimize_compound_coords(reference_compound_coords, predict_compound_coords, total_epoch=1000, LAS_edge_index=None, mode=0):
    """Same signature and return triple as post_optim_utils.py:36-64 for ONE ligand: (x [n,3], last loss, last rmsd).
    `mode` only selects the unused interaction-loss variant in the reference (post_optim_utils.py:13-21,33-34)."""
    n = predict_compound_coords.shape[0]
    dev = predict_compound_coords.device
    batch = torch.zeros(n, dtype=torch.long, device=dev)
    las_b = None if LAS_edge_index is None else torch.zeros(LAS_edge_index.shape[1], dtype=torch.long, device=dev)
    x, loss, rmsd = post_optimize_batch(reference_compound_coords, predict_compound_coords, batch, LAS_edge_index, las_b, total_epoch)
    return x, float(loss[0]), float(rmsd[0])
