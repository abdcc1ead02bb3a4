"""Reverse-pass primitives and the MC_E_GCL reverse pass on the GPU (csrc/backward.cu through the C ABI, fabind_b200/backward.py)
against torch autograd of the same formulas (fp32 reference of the op, tolerance 1e-4 of the tensor's scale: scatter directions
use fp32 atomics).  The formulas are those of tests/emulate_backward.py, which is pinned against the unmodified reference."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _graph(B, n_lo, n_hi, deg, seed):
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(n_lo, n_hi, (B,), generator=g).tolist()
    row, col, cplx = [], [], []
    o = 0
    for b, n in enumerate(sizes):
        cplx += [b] * n
        for i in range(n):
            if i == 1:
                continue                                   # an isolated node: degree 0 -> count clamps to 1
            k = deg if i else min(n - 1, 4 * deg)          # node 0 plays the high-degree global node
            nb = torch.randperm(n - 1, generator=g)[:k]
            nb = nb + (nb >= i).long()
            row += [o + i] * len(nb)
            col += (o + nb).tolist()
        o += n
    return torch.tensor(row), torch.tensor(col), torch.tensor(cplx), o, B


def gcl_reference(p, h, x, row, col, cplx, B, cmax):
    """torch restatement of the launch sequence of one MC_E_GCL (tests/emulate_backward.py::gcl_fwd), any device"""
    N, H = h.shape
    d = x[row] - x[col]
    d2 = (d * d).sum(1)
    nrm = torch.zeros(B, device=h.device).index_add_(0, cplx[row], d2 * d2).sqrt()
    rn = d2 / nrm[cplx[row]]
    Pn = F.linear(h, p["e1_rc"])
    Z1 = Pn[row, :H] + Pn[col, H:] + rn[:, None] * p["e1_rad"] + p["e1_b"]
    Z2 = F.linear(F.silu(Z1), p["e2_w"], p["e2_b"])
    M = F.silu(Z2)
    Z3 = F.linear(M, p["c1_w"], p["c1_b"])
    s = F.silu(Z3) @ p["c2_w"]
    deg = torch.zeros(N, device=h.device).index_add_(0, row, torch.ones(row.numel(), device=h.device)).clamp(min=1)
    step = torch.zeros(N, 3, device=h.device).index_add_(0, row, d * s[:, None]) / deg[:, None]
    x_new = x + step.clamp(-cmax, cmax)
    agg = torch.zeros(N, H, device=h.device).index_add_(0, row, M)
    Z4 = F.linear(torch.cat([h, agg], 1), p["n1_w"], p["n1_b"])
    h_new = h + F.linear(F.silu(Z4), p["n2_w"], p["n2_b"])
    saved = dict(h=h, x=x, rn=rn, nrm=nrm, Z1=Z1, Z2=Z2, Z3=Z3, s=s, deg=deg, step=step, agg=agg, Z4=Z4)
    return h_new, x_new, saved


def gcl_problem(H, seed, device):
    row, col, cplx, N, B = _graph(3, 30, 60, 6, seed)
    g = torch.Generator().manual_seed(seed + 1)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc)
    p = dict(e1_rc=r(2 * H, H, sc=H ** -0.5), e1_rad=r(H), e1_b=r(H, sc=0.1), e2_w=r(H, H, sc=H ** -0.5), e2_b=r(H, sc=0.1),
             c1_w=r(H, H, sc=H ** -0.5), c1_b=r(H, sc=0.1), c2_w=r(H, sc=2.0 * H ** -0.5), n1_w=r(H, 2 * H, sc=(2 * H) ** -0.5),
             n1_b=r(H, sc=0.1), n2_w=r(H, H, sc=H ** -0.5), n2_b=r(H, sc=0.1))
    h, x = r(N, H, sc=0.5), r(N, 3, sc=1.5)
    gh, gx = r(N, H), r(N, 3)
    mv = lambda t: t.to(device)
    return ({k: mv(v) for k, v in p.items()}, mv(h), mv(x), mv(row), mv(col), mv(cplx), B, mv(gh), mv(gx))


def test_gcl_reference_exercises_the_clamp():
    """the test problem has clamped and unclamped coordinate steps, so the reverse pass of the clamp is exercised"""
    p, h, x, row, col, cplx, B, gh, gx = gcl_problem(64, 5, "cpu")
    _, _, sv = gcl_reference(p, h, x, row, col, cplx, B, 0.5)
    frac = float((sv["step"].abs() > 0.5).float().mean())
    assert 0.05 < frac < 0.95, frac


@pytest.mark.parametrize("H", [64, 128])
def test_gcl_backward_matches_autograd(H):
    from fabind_b200 import backward as bw
    dev = "cuda"
    cmax = 0.5
    p, h, x, row, col, cplx, B, gh, gx = gcl_problem(H, 5, dev)
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    hl, xl = h.clone().requires_grad_(True), x.clone().requires_grad_(True)
    h_new, x_new, sv = gcl_reference(leaves, hl, xl, row, col, cplx, B, cmax)
    ((h_new * gh).sum() + (x_new * gx).sum()).backward()
    saved = {k: v.detach().contiguous() for k, v in sv.items()}
    w = {k: v.contiguous() for k, v in p.items()}
    for k in ("e1_rc", "e2_w", "c1_w", "n1_w", "n2_w"):
        w[k + "_t"] = p[k].t().contiguous()
    dh, dx, grads = bw.gcl_backward(w, saved, row.int().contiguous(), col.int().contiguous(), cplx.int().contiguous(), cmax,
                                    gh.contiguous(), gx.contiguous())
    torch.cuda.synchronize()
    assert rel_err(dh, hl.grad) < TOL, ("dh", rel_err(dh, hl.grad))
    assert rel_err(dx, xl.grad) < TOL, ("dx", rel_err(dx, xl.grad))
    assert set(grads) == set(p), set(grads) ^ set(p)
    for k, v in grads.items():
        assert rel_err(v, leaves[k].grad) < TOL, (k, rel_err(v, leaves[k].grad))


def test_primitives_match_torch():
    from fabind_b200 import backward as bw
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    Z, dY = torch.randn(1000, 70, generator=g).to(dev), torch.randn(1000, 70, generator=g).to(dev)
    for act, fn in ((bw.ACT_SILU, F.silu), (bw.ACT_RELU, F.relu)):
        zl = Z.clone().requires_grad_(True)
        y = fn(zl)
        y.backward(dY)
        assert rel_err(bw.act_fwd(Z, act), y.detach()) < 1e-6
        assert rel_err(bw.act_bwd(Z, dY, act), zl.grad) < 1e-5
    u, v = torch.randn(1000, generator=g).to(dev), torch.randn(70, generator=g).to(dev)
    zl = Z.clone().requires_grad_(True)
    ((F.silu(zl) @ v) * u).sum().backward()
    assert rel_err(bw.outer_act_bwd(Z, u, v, bw.ACT_SILU), zl.grad) < 1e-5
    assert rel_err(bw.colsum(Z), Z.sum(0)) < 1e-5
    assert rel_err(bw.colsum(Z, u), (Z * u[:, None]).sum(0)) < 1e-5
    assert rel_err(bw.rowdot(Z, v), Z @ v) < 1e-5
    X = torch.randn(1000, 45, generator=g).to(dev)
    assert rel_err(bw.gemm_wgrad(dY, X), dY.t() @ X) < 1e-5
    big_dY, big_X = torch.randn(20000, 128, generator=g).to(dev), torch.randn(20000, 256, generator=g).to(dev)
    assert rel_err(bw.gemm_wgrad(big_dY, big_X), big_dY.t() @ big_X) < 2e-5
    dY2, Wt = torch.randn(1000, 72, generator=g).to(dev), torch.randn(48, 72, generator=g).to(dev)   # W^T of a Linear(48 -> 72)
    assert rel_err(bw.gemm_dgrad(dY2, Wt), dY2 @ Wt.t()) < 1e-5
    idx = torch.randint(0, 50, (1000,), generator=g).to(dev)
    dst = torch.zeros(50, 100, device=dev)
    bw.scatter_add_rows(Z, idx.int(), dst, col0=20)
    ref = torch.zeros(50, 100, device=dev)
    ref[:, 20:90].index_add_(0, idx, Z)
    assert rel_err(dst, ref) < 1e-5
    src = torch.randn(50, 100, generator=g).to(dev)
    out = Z.clone()
    bw.gather_add_rows(src, idx.int(), out, col0=20)
    assert rel_err(out, Z + src[idx, 20:90]) < 1e-6


def test_las_backward_matches_autograd():
    from fabind_b200 import backward as bw
    dev = "cuda"
    g = torch.Generator().manual_seed(9)
    N, E = 60, 400
    x = (torch.randn(N, 3, generator=g) * 1.2).to(dev)
    xref = (x.cpu() + 0.3 * torch.randn(N, 3, generator=g)).to(dev)
    a, b = torch.randint(0, N, (E,), generator=g).to(dev), torch.randint(0, N, (E,), generator=g).to(dev)
    gx = torch.randn(N, 3, generator=g).to(dev)
    step_size, lcl = 0.02, 0.6
    xl = x.clone().requires_grad_(True)
    d = xl[a] - xl[b]
    diff = (d * d).sum(1) - ((xref[a] - xref[b]) ** 2).sum(1)
    acc = torch.zeros_like(x).index_add_(0, b, 4 * diff[:, None] * d) * step_size
    frac = float((acc.abs() > lcl).float().mean())
    assert 0.02 < frac < 0.98, frac
    ((xl + acc.clamp(-lcl, lcl)) * gx).sum().backward()
    dx = bw.las_bwd(x.contiguous(), xref.contiguous(), a.int().contiguous(), b.int().contiguous(), acc.detach().contiguous(),
                    step_size, lcl, gx.contiguous())
    assert rel_err(dx, xl.grad) < TOL, rel_err(dx, xl.grad)


def test_backward_refuses_cpu_tensors():
    from fabind_b200 import backward as bw
    with pytest.raises(RuntimeError):
        bw.act_fwd(torch.zeros(4, 4), bw.ACT_SILU)
