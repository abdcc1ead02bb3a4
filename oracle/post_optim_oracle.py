"""CPU oracle for the ligand post-optimisation (TEST INFRASTRUCTURE: the checker, never the product).

Restates `post_optimize_compound_coords` / `post_optimize_loss_function` of FABind/fabind/utils/post_optim_utils.py:8-64 with
the gradient written out (what autograd computes there) and torch.optim.Adam's update rule (lr 0.1, betas (0.9, 0.999),
eps 1e-8, no weight decay).  Only `configuration_loss` enters the objective (post_optim_utils.py:33-34); `mode` changes the
reported interaction loss only, which is not part of the outputs used by callers (fabind_inference.py:295-316).
Pinned against the reference function itself (tests/golden/postopt_*.pt, scripts/make_golden.py::main_post_optim).
"""
import torch


def _loss_and_grad(x, c, mask):
    """loss = sum_{mask} |d_ij - c_ij| + 2 sum_{all i,j} relu(1.22 - d_ij)   (mask given: post_optim_utils.py:25-29)
            = sum_{all i,j} |d_ij - c_ij|                                      (mask None:  post_optim_utils.py:31)
    over ORDERED pairs of the full n x n matrix; torch.cdist's backward gives 0 where d == 0, abs'(0) = 0, relu'(0) = 0."""
    diff = x[:, None, :] - x[None, :, :]
    d = diff.pow(2).sum(-1).sqrt()
    err = d - c
    if mask is not None:
        loss = err.abs()[mask].sum() + 2 * (1.22 - d).relu().sum()
        w = torch.sign(err) * mask.to(x.dtype) - 2.0 * (d < 1.22).to(x.dtype)
    else:
        loss = err.abs().sum()
        w = torch.sign(err)
    unit = torch.where(d[..., None] > 0, diff / d.clamp_min(1e-30)[..., None], torch.zeros_like(diff))
    g = (w[..., None] * unit).sum(1) - (w[..., None] * unit).sum(0)      # pair (i,j) touches x_i (+) and x_j (-)
    return loss, g


def post_optimize(reference_coords, predict_coords, total_epoch=1000, las_edge_index=None, lr=0.1):
    """-> (x [n,3], last loss (before the last step), rmsd to the reference coordinates after the last step)"""
    n = predict_coords.shape[0]
    mask = None
    if las_edge_index is not None:
        mask = torch.zeros((n, n), dtype=torch.bool)
        mask[las_edge_index[0], las_edge_index[1]] = True               # to_dense_adj(LAS_edge_index)
    c = torch.cdist(reference_coords, reference_coords)
    x = predict_coords.clone()
    m = torch.zeros_like(x)
    v = torch.zeros_like(x)
    b1, b2, eps = 0.9, 0.999, 1e-8
    loss = torch.zeros(())
    for t in range(1, total_epoch + 1):
        loss, g = _loss_and_grad(x, c, mask)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        denom = v.sqrt() / (1 - b2 ** t) ** 0.5 + eps
        x = x - (lr / (1 - b1 ** t)) * m / denom
    rmsd = ((reference_coords - x) ** 2).sum(-1).mean().sqrt()
    return x, float(loss), float(rmsd)
