"""Training-mode forward of the last refinement iteration (fabind_b200/backward.py::stack_forward_train_v1) and the closed loop
forward -> reverse on the REAL kernels, against the pinned specification; the assembled training step (fabind_b200/train.py) against
parameter gradients of the unmodified reference, without and with training-mode dropout; the drop-in module in train() mode.
(First run on a B200 in round 2: 10/10 green, profiles/r2a_train_forward_tests.txt.)"""
import glob
import os

import pytest
import torch

from helpers import GOLDEN_DIR, rel_err

pytestmark = [pytest.mark.gpu]
TOL = 1e-4


def _cuda(o):
    if torch.is_tensor(o):
        return o.cuda().contiguous()
    if isinstance(o, dict):
        return {k: _cuda(v) for k, v in o.items()}
    return o


def _cpu(o):
    if torch.is_tensor(o):
        return o.cpu()
    if isinstance(o, dict):
        return {k: _cpu(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return type(o)(_cpu(v) for v in o)
    return o


def test_forward_pieces_match_torch():
    from fabind_b200 import backward as bw
    g = torch.Generator().manual_seed(5)
    N, E, B = 200, 3000, 4
    x = torch.randn(N, 3, generator=g)
    cplx = torch.sort(torch.randint(0, B, (N,), generator=g)).values
    row = torch.sort(torch.randint(0, N, (E,), generator=g)).values
    col = torch.randint(0, N, (E,), generator=g)
    d = x[row] - x[col]
    d2 = (d * d).sum(1)
    nrm = torch.zeros(B).index_add_(0, cplx[row], d2 * d2).sqrt()
    gd, gd2, grn, gnrm = bw.radial_fwd(x.cuda(), row.int().cuda(), col.int().cuda(), cplx.int().cuda(), B)
    assert rel_err(gd, d) < 1e-6 and rel_err(gd2, d2) < 1e-6 and rel_err(gnrm, nrm) < 1e-5 and rel_err(grn, d2 / nrm[cplx[row]]) < 1e-5
    ssum, cnt = torch.randn(N, 3, generator=g) * 3, torch.randint(0, 5, (N,), generator=g).float()
    st, xn = bw.coord_apply(x.cuda(), ssum.cuda(), cnt.cuda(), 1.0)
    ref = ssum / cnt.clamp(min=1)[:, None]
    assert rel_err(st, ref) < 1e-6 and rel_err(xn, x + ref.clamp(-1, 1)) < 1e-6
    logit = torch.randn(E, generator=g) * 3
    rowptr = torch.zeros(N + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=N), 0)
    mx = torch.full((N,), float("-inf")).scatter_reduce(0, row, logit, reduce="amax")
    e = (logit - mx[row]).exp()
    assert rel_err(bw.softmax_seg_fwd(logit.cuda(), rowptr.int().cuda(), N), e / torch.zeros(N).index_add_(0, row, e)[row]) < 1e-5
    xref = x + 0.3 * torch.randn(N, 3, generator=g)
    a, b = torch.randint(0, N, (500,), generator=g), torch.randint(0, N, (500,), generator=g)
    dd = x[a] - x[b]
    diff = (dd * dd).sum(1) - ((xref[a] - xref[b]) ** 2).sum(1)
    assert rel_err(bw.las_acc(x.cuda(), xref.cuda(), a.int().cuda(), b.int().cuda(), 0.02),
                   torch.zeros_like(x).index_add_(0, b, 4 * diff[:, None] * dd * 0.02)) < 1e-5
    raw = torch.randn(1000, 128, generator=g)
    r5 = raw[:, :64].reshape(1000, 8, 2, 4)
    assert rel_err(bw.pair_bias_gate_fwd(raw.cuda(), 8), r5[:, :, 0] * torch.sigmoid(r5[:, :, 1])) < 1e-5


def test_training_forward_and_reverse_on_the_real_kernels():
    from fabind_b200 import backward as bw
    from test_backward_orchestration import stack_case, check_forward, check_stack, two_layer_problem
    cases = [stack_case(p) for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))] + [stack_case(problem=two_layer_problem())]
    for case in cases:
        w, geo, edges, consts = _cuda(case["weights"]), _cuda(case["geo"]), _cuda(case["edges"]), _cuda(case["consts"])
        X_out, H_out, tape, top = bw.stack_forward_train_v1(w, case["top"]["Hin"].cuda(), case["x_state"].cuda(), case["moves"].cuda(), geo,
                                                            edges, consts, case["L"])
        torch.cuda.synchronize()
        check_forward(case, X_out.cpu(), H_out.cpu(), _cpu(tape), _cpu(top), TOL)
        grads, dHin = bw.stack_backward_v1(w, tape, top, geo, edges, consts, case["dH_out"].cuda(), case["dX_out"].cuda())
        torch.cuda.synchronize()
        check_stack(case, _cpu(grads), dHin.cpu(), TOL)


def test_training_step_on_the_gpu():
    """fabind_b200/train.py::training_step_v1 with the real providers (earlier iterations through fb_model_forward, edge lists from
    the graph builder) and the real kernels: parameter gradients of the unmodified reference (tests/golden/grad_v1_*.pt)"""
    from fabind_b200 import EfficientMCAttModel, train
    from fabind_b200.config import published_args
    from helpers import load_golden
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt"))):
        g, r, b, sd, cfg = load_golden(path)
        H = r["hidden"]
        model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                                    normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen).cuda(), torch.randn(b.H.shape, generator=gen).cuda()
        fa = b.to("cuda").forward_args()
        X_out, H_out, pgrads, gH_in = train.training_step_v1(model, fa, lambda X, Hh: (rx, rh))
        torch.cuda.synchronize()
        loss = float((X_out * rx).sum() + (H_out * rh).sum())
        assert abs(loss - g["loss"]) < 1e-3 * abs(g["loss"]), (loss, g["loss"])
        gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
        n = 0
        for k, ref in g["grads"].items():
            if ref is None:
                continue
            err = float((pgrads[k].cpu() - ref).abs().max())
            assert err < 1e-3 * float(ref.abs().max()) + 1e-5 * gmax, (k, err, float(ref.abs().max()))
            n += 1
        assert n >= 80


def test_plus_kernels_match_torch():
    from fabind_b200 import backward as bw
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(6)
    M, D = 900, 160
    x = (torch.randn(M, D, generator=g) * 2 + 0.5).requires_grad_(True)
    gamma, beta, dy = torch.randn(D, generator=g), torch.randn(D, generator=g), torch.randn(M, D, generator=g)
    y = F.layer_norm(x, (D,), gamma.clone().requires_grad_(True), beta, 1e-5)
    gl = gamma.clone().requires_grad_(True)
    bl = beta.clone().requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    F.layer_norm(x2, (D,), gl, bl, 1e-5).backward(dy)
    grads = {}
    dx = bw.layernorm_bwd(grads, "g", "b", x.detach().cuda(), gamma.cuda(), dy.cuda())
    assert rel_err(dx, x2.grad) < 1e-5 and rel_err(grads["g"], gl.grad) < 1e-5 and rel_err(grads["b"], bl.grad) < 1e-5
    h, w = torch.randn(M, D, generator=g), torch.randn(D, generator=g)
    ds1, ds2, ds3 = torch.randn(M, generator=g), torch.randn(M, generator=g), torch.randn(M, generator=g)
    base = torch.randn(M, D, generator=g)
    out = bw.row_stats_bwd(h.cuda(), w.cuda(), ds1.cuda(), ds2.cuda(), ds3.cuda(), base.clone().cuda())
    assert rel_err(out, base + ds1[:, None] + 2 * h * ds2[:, None] + w[None, :] * ds3[:, None]) < 1e-6
    E, Dn = 4000, 65.0
    A1, A2, A3, rn = (torch.randn(E, generator=g).requires_grad_(True), (torch.rand(E, generator=g) * 50 + 5).requires_grad_(True),
                      torch.randn(E, generator=g).requires_grad_(True), torch.rand(E, generator=g).requires_grad_(True))
    a = torch.tensor([0.7, 1.9], requires_grad=True)
    mu = (A1 + rn * a[0]) / Dn
    var_raw = (A2 + 2 * rn * A3 + rn * rn * a[1]) / Dn - mu * mu
    rstd = torch.rsqrt(var_raw.clamp(min=0) + 1e-5)
    drstd, dmu = torch.randn(E, generator=g), torch.randn(E, generator=g)
    ((rstd * drstd).sum() + (mu * dmu).sum()).backward()
    drn = torch.zeros(E).cuda()
    c = lambda t: t.detach().cuda()
    dA1, dA2, dA3, da = bw.folded_stats_bwd(c(A3), c(rn), 0.7, 1.9, Dn, c(mu), c(var_raw), c(rstd), drstd.cuda(), dmu.cuda(), drn, True)
    assert rel_err(dA1, A1.grad) < 1e-5 and rel_err(dA2, A2.grad) < 1e-5 and rel_err(dA3, A3.grad) < 1e-5
    assert rel_err(drn, rn.grad) < 1e-5 and rel_err(da, a.grad) < 1e-4


def test_plus_stack_backward_on_the_real_kernels():
    from fabind_b200 import backward as bw
    from test_backward_orchestration import plus_stack_case, check_stack
    case = plus_stack_case()
    tape = [tuple(_cuda(s) for s in layer) for layer in case["tape"]]
    grads, dHin = bw.stack_backward_plus(_cuda(case["weights"]), tape, _cuda(case["top"]), _cuda(case["geo"]), _cuda(case["edges"]),
                                         _cuda(case["consts"]), case["dH_out"].cuda(), case["dX_out"].cuda(), case["dP_out"].cuda())
    torch.cuda.synchronize()
    check_stack(case, _cpu(grads), dHin.cpu(), TOL)


def test_plus_training_forward_and_reverse_on_the_real_kernels():
    from fabind_b200 import backward as bw
    from test_backward_orchestration import plus_stack_case, check_stack
    case = plus_stack_case()
    case["consts"]["n_pairs"] = int(case["dP_out"].shape[0])
    w, geo, edges, consts = _cuda(case["weights"]), _cuda(case["geo"]), _cuda(case["edges"]), _cuda(case["consts"])
    X_out, H_out, pair, tape, top = bw.stack_forward_train_plus(w, case["top"]["Hin"].cuda(), case["x_state"].cuda(), case["moves"].cuda(), geo,
                                                                edges, consts, len(case["tape"]))
    torch.cuda.synchronize()
    for mine, ref in zip(_cpu(tape) + [(_cpu(top["out_saved"]),)], case["tape"] + [(case["top"]["out_saved"],)]):
        for sm, sr in zip(mine, ref):
            for k in sr:
                if k == "acr":
                    continue
                if sr[k].dtype == torch.int32:
                    assert torch.equal(sm[k], sr[k]), k
                else:
                    assert rel_err(sm[k], sr[k]) < TOL, (k, rel_err(sm[k], sr[k]))
    grads, dHin = bw.stack_backward_plus(w, tape, top, geo, edges, consts, case["dH_out"].cuda(), case["dX_out"].cuda(), case["dP_out"].cuda())
    torch.cuda.synchronize()
    check_stack(case, _cpu(grads), dHin.cpu(), TOL)


def test_plus_training_step_on_the_gpu():
    """train.training_step on the FABind+ layout with the real providers and kernels: parameter gradients of the unmodified FABind+
    reference (tests/golden/grad_plus_*.pt)"""
    from fabind_b200 import train
    from fabind_b200.config import published_args_plus
    from fabind_b200.plus import EfficientMCAttModel as PlusModel
    from helpers import load_golden
    from test_formulation_cpu import _dense_pair
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_plus_*.pt"))):
        g, r, b, sd, cfg = load_golden(path)
        H = r["hidden"]
        model = PlusModel(published_args_plus(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                          normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        model.load_state_dict(sd, strict=True)
        model = model.cuda().eval()
        model.return_pair = False
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen), torch.randn(b.H.shape, generator=gen)
        dims = [(int(b.n_p[i]) + 1, int(b.n_c[i]) + 1) for i in range(len(b.n_c))]
        seen = {}

        def output_grads(X, Hh, pair):
            dense = _dense_pair(pair.cpu(), dims, H)
            rp = torch.randn(dense.shape, generator=gen) * 0.1
            seen["loss"] = float((X.cpu() * rx).sum() + (Hh.cpu() * rh).sum() + (dense * rp).sum())
            return rx.cuda(), rh.cuda(), torch.cat([rp[i, :n, :c].reshape(-1, H) for i, (n, c) in enumerate(dims)]).cuda()
        out = train.training_step(model, b.to("cuda").forward_args(), output_grads)
        torch.cuda.synchronize()
        pgrads = out[3]
        assert abs(seen["loss"] - g["loss"]) < 1e-3 * abs(g["loss"]), (seen["loss"], g["loss"])
        gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
        n = 0
        for k, ref in g["grads"].items():
            if ref is None:
                continue
            err = float((pgrads[k].cpu() - ref).abs().max())
            assert err < 1e-3 * float(ref.abs().max()) + 1e-5 * gmax, (k, err, float(ref.abs().max()))
            n += 1
        assert n >= 80


def test_forward_with_grad_on_the_gpu():
    """an unchanged training loop on the drop-in module: loss.backward() through train.forward_with_grad puts the unmodified
    reference's gradients on the module's parameters (real providers, real kernels)"""
    from fabind_b200 import EfficientMCAttModel, train
    from fabind_b200.config import published_args
    from helpers import load_golden
    g, r, b, sd, cfg = load_golden(sorted(glob.glob(os.path.join(GOLDEN_DIR, "grad_v1_*.pt")))[0])
    H = r["hidden"]
    model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                                normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    gen = torch.Generator().manual_seed(r["readout_seed"])
    rx, rh = torch.randn(b.X.shape, generator=gen).cuda(), torch.randn(b.H.shape, generator=gen).cuda()
    fa = b.to("cuda").forward_args()
    X, Hh = train.forward_with_grad(model, fa, n_iter=r["n_iter"])     # (train() mode would draw randint(1, n_iter): pin it for the golden)
    loss = (X * rx).sum() + (Hh * rh).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - g["loss"]) < 1e-3 * abs(g["loss"])
    params = dict(model.named_parameters())
    gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
    for k, ref in g["grads"].items():
        if ref is not None:
            err = float((params[k].grad.cpu() - ref).abs().max())
            assert err < 1e-3 * float(ref.abs().max()) + 1e-5 * gmax, (k, err)


def _check_grads(pgrads, g, rel=1e-3, floor=1e-5):
    gmax = max(float(v.abs().max()) for v in g["grads"].values() if v is not None)
    n = 0
    for k, ref in g["grads"].items():
        if ref is None:
            continue
        got = pgrads[k]
        got = got.grad if isinstance(got, torch.nn.Parameter) else got
        err = float((got.cpu() - ref).abs().max())
        assert err < rel * float(ref.abs().max()) + floor * gmax, (k, err, float(ref.abs().max()))
        n += 1
    assert n >= 80


def _dropout_model(r, sd):
    from fabind_b200 import EfficientMCAttModel
    from fabind_b200.config import published_args
    H = r["hidden"]
    args = published_args()
    args.random_n_iter = False
    model = EfficientMCAttModel(args, H, H, 1, n_layers=r["n_layers"], dropout=r["dropout_p"], n_iter=r["n_iter"],
                                normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    model.dropout_seed, model.dropout_colonly = r["dropout_seed"], True
    return model


def test_train_mode_forward_with_dropout_matches_the_reference():
    """train() under no_grad: every refinement iteration through the inference kernels with the v1 dropout sites active
    (egnn.py:82,106,236,398,461; cross_att.py:128) -- outputs of the unmodified reference in train() mode whose nn.Dropout modules
    were patched to the library's column-only masks (tests/golden/graddrop_v1_*.pt)"""
    from helpers import load_golden
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, "graddrop_v1_*.pt")))
    assert paths
    for path in paths:
        g, r, b, sd, cfg = load_golden(path)
        model = _dropout_model(r, sd)
        for prec, tol in (("fp32", TOL), ("fp32_tc", TOL)):
            model.precision = prec
            bc = b.to("cuda")
            with torch.no_grad():
                X, Hh = model(**bc.forward_args())
            torch.cuda.synchronize()
            assert rel_err(X.cpu(), g["X"]) < tol and rel_err(Hh.cpu(), g["H"]) < tol, (path, prec, rel_err(X.cpu(), g["X"]), rel_err(Hh.cpu(), g["H"]))


def test_training_step_with_dropout_on_the_gpu():
    """the drop-in module in train() mode with autograd: forward() routes to train.forward_with_grad, dropout masks in the no_grad
    iterations, the training-mode forward and the reverse pass; loss.backward() leaves the unmodified reference's train()-mode
    gradients on the parameters"""
    from helpers import load_golden
    for path in sorted(glob.glob(os.path.join(GOLDEN_DIR, "graddrop_v1_*.pt"))):
        g, r, b, sd, cfg = load_golden(path)
        model = _dropout_model(r, sd)
        gen = torch.Generator().manual_seed(r["readout_seed"])
        rx, rh = torch.randn(b.X.shape, generator=gen).cuda(), torch.randn(b.H.shape, generator=gen).cuda()
        bc = b.to("cuda")
        X, Hh = model(**bc.forward_args())
        loss = (X * rx).sum() + (Hh * rh).sum()
        loss.backward()
        torch.cuda.synchronize()
        assert rel_err(X.detach().cpu(), g["X"]) < TOL and rel_err(Hh.detach().cpu(), g["H"]) < TOL
        assert abs(float(loss.detach()) - g["loss"]) < 1e-3 * abs(g["loss"])
        _check_grads(dict(model.named_parameters()), g)
        assert model.training and all(m.training for m in model.modules())      # the step does not flip module modes


def test_row_column_masks_match_the_specification(monkeypatch):
    """full row x column masks (what training uses; no reference can reproduce them): the training step on the real kernels against
    the SAME orchestration with every kernel wrapper replaced by its torch definition and the masks taken from the numpy restatement
    of the hash (fabind_b200/dropout.py) -- pins the row / column indexing of fb_dropout_apply and of the GEMM-epilogue masks in the
    forward and in the reverse pass (placement and scaling are pinned by the column-only goldens above)"""
    from fabind_b200 import EfficientMCAttModel, backward as bw, train
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_batch, randomize_coord_heads
    from oracle import fabind_oracle as orc
    import test_backward_orchestration as tbo
    H, L = 64, 2
    torch.manual_seed(1)
    args = published_args()
    args.random_n_iter = False
    model = EfficientMCAttModel(args, H, H, 1, n_layers=L, dropout=0.2, n_iter=1, normalize_coord=lambda x: x / 5.0,
                                unnormalize_coord=lambda x: x * 5.0)
    randomize_coord_heads(model, std=0.3)
    b = make_batch(n_complexes=2, seed=4, embed=H, n_c_range=(8, 12), n_p_range=(30, 40))
    g = torch.Generator().manual_seed(3)
    rx, rh = torch.randn(b.X.shape, generator=g), torch.randn(b.H.shape, generator=g)
    dropout = (0.2, 99, False)
    model = model.cuda().train()
    out = train.training_step(model, b.to("cuda").forward_args(), lambda X, Hh: (rx.cuda(), rh.cuda()), dropout=dropout)
    torch.cuda.synchronize()
    Xg, Hg, pg = out[0].cpu(), out[1].cpu(), {k: v.cpu() for k, v in out[2].items()}
    # the masks are not trivially off: the dropped step differs from the undropped one (run before the stand-ins are installed)
    plain = train.training_step(model, b.to("cuda").forward_args(), lambda X, Hh: (rx.cuda(), rh.cuda()))
    assert rel_err(plain[1].cpu(), Hg) > 1e-2
    # the same step with torch stand-ins for every kernel wrapper, on the CPU
    tbo._install_standins(monkeypatch, bw)
    tbo._install_forward_standins(monkeypatch, bw)
    monkeypatch.setattr(bw, "pair_bias_gate_bwd", tbo._gate_bwd_standin)
    monkeypatch.setattr(bw, "pair_outer_bwd", tbo._outer_bwd_standin)
    cpu = model.cpu()
    cfg = orc.make_cfg(n_layers=L, n_iter=1)

    def edge_lists(m, X_prev, fa):
        ctx, inter, _ = orc.build_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"], cfg.intra_cutoff / cfg.coordinate_scale,
                                        cfg.inter_cutoff / cfg.coordinate_scale)
        return ctx, inter
    ref = train.training_step(cpu, b.forward_args(), lambda X, Hh: (rx, rh), prev_coords=lambda m, fa: fa["X"].clone(), edge_lists=edge_lists,
                              dropout=dropout)
    assert rel_err(Xg, ref[0]) < TOL and rel_err(Hg, ref[1]) < TOL, (rel_err(Xg, ref[0]), rel_err(Hg, ref[1]))
    gmax = max(float(v.abs().max()) for v in ref[2].values())
    for k, r in ref[2].items():
        err = float((pg[k] - r).abs().max())
        assert err < 1e-3 * float(r.abs().max()) + 1e-5 * gmax, (k, err, float(r.abs().max()))


def test_training_step_bf16_gemms_close_to_fp32():
    """PRECISION = 'bf16': forward / data-gradient GEMMs on the tcgen05 kernels, weight gradients as a tcgen05 GEMM over transposed
    bf16 copies; gradients stay within bf16 distance of the fp32 parity path (the reference has no bf16 mode: own bound, 5 % of scale)"""
    from fabind_b200 import EfficientMCAttModel, backward as bw, train
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_batch
    H, L = 128, 2
    torch.manual_seed(0)
    model = EfficientMCAttModel(published_args(), H, H, 1, n_layers=L, n_iter=2, normalize_coord=lambda x: x / 5.0,
                                unnormalize_coord=lambda x: x * 5.0).cuda().eval()
    b = make_batch(n_complexes=4, seed=2, embed=H, n_c_range=(20, 40), n_p_range=(120, 200)).to("cuda")
    g = torch.Generator().manual_seed(3)
    rx, rh = torch.randn(b.X.shape, generator=g).cuda(), torch.randn(b.H.shape, generator=g).cuda()
    res = {}
    X0 = b.X.clone()
    for prec in ("fp32", "bf16"):
        bw.PRECISION = prec
        try:
            fa = b.forward_args()
            fa["X"] = X0.clone()
            res[prec] = train.training_step(model, fa, lambda X, Hh: (rx, rh))[2]
        finally:
            bw.PRECISION = "fp32"
    gmax = max(float(v.abs().max()) for v in res["fp32"].values())
    for k, ref in res["fp32"].items():
        err = float((res["bf16"][k] - ref).abs().max())
        assert err < 5e-2 * float(ref.abs().max()) + 5e-3 * gmax, (k, err, float(ref.abs().max()))


def test_row_attention_reverse_long_key_lists():
    """more than 256 keys (whole proteins of the pocket stage): the global-memory variant of the row-attention reverse"""
    import emulate_backward as spec
    from fabind_b200 import backward as bw
    g = torch.Generator().manual_seed(8)
    nc1, np1 = 21, 301
    geo = dict(Nc=nc1, B=1, max_c=nc1, max_p=np1, c_off=torch.tensor([0, nc1], dtype=torch.int32), p_off=torch.tensor([nc1, nc1 + np1], dtype=torch.int32),
               pair_base=torch.tensor([0, nc1 * np1], dtype=torch.int32), node_cplx=torch.zeros(nc1 + np1, dtype=torch.int32))
    CAc, CAp2 = torch.randn(nc1, 512, generator=g), torch.randn(np1, 256, generator=g)
    PB, dO = torch.randn(np1 * nc1, 4, generator=g), torch.randn(nc1, 128, generator=g)
    _, sv = spec.rowatt_fwd(CAc[:, 256:384], CAc[:, 384:], CAp2[:, :128], CAp2[:, 128:], PB.view(np1, nc1, 4).transpose(0, 1))
    dq, dg, dk, dv, db = spec.rowatt_bwd(sv, dO)
    gd = _cuda(geo)
    dCAc, dCAp2 = torch.zeros_like(CAc).cuda(), torch.zeros_like(CAp2).cuda()
    dPB = bw.row_attention_bwd(gd, 0, (CAc.cuda(), 256), (CAc.cuda(), 384), (CAp2.cuda(), 0), (CAp2.cuda(), 128), PB.cuda(), dO.cuda(),
                               (dCAc, 256), (dCAc, 384), (dCAp2, 0), (dCAp2, 128))
    torch.cuda.synchronize()
    assert rel_err(dCAc[:, 256:384], dq) < TOL and rel_err(dCAc[:, 384:], dg) < TOL
    assert rel_err(dCAp2[:, :128], dk) < TOL and rel_err(dCAp2[:, 128:], dv) < TOL
    assert rel_err(dPB, db.transpose(0, 1).reshape(-1, 4)) < TOL
