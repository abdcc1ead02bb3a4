"""Fused linear layer (fb_gemm) against torch on the same device, through the C ABI."""
import ctypes as C

import pytest
import torch

from fabind_b200 import _lib

pytestmark = pytest.mark.gpu


def run_gemm(A, W, bias=None, act=0, res=None, A2=None, dotv=None, m_dev=None, bf16=False, want_cb=False,
             force_simt=False, want_c=True):
    l = _lib.lib()
    dev = A.device
    M, K1 = A.shape
    K2 = A2.shape[1] if A2 is not None else 0
    N = W.shape[0]
    g = _lib.GemmParams()
    dt = torch.bfloat16 if bf16 else torch.float32
    Ad, Wd = A.to(dt).contiguous(), W.to(dt).contiguous()
    A2d = A2.to(dt).contiguous() if A2 is not None else None
    g.A, g.lda, g.K1 = Ad.data_ptr(), K1, K1
    g.A2, g.lda2, g.K2 = (A2d.data_ptr() if A2d is not None else None), K2, K2
    g.W = Wd.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.act = act
    g.res, g.ldres = (res.data_ptr() if res is not None else None), N
    Cout = torch.full((M, N), float("nan"), device=dev)
    g.C, g.ldc = (Cout.data_ptr() if want_c else None), N
    Cb = torch.zeros((M, N), dtype=dt, device=dev) if want_cb else None
    g.Cb, g.ldcb = (Cb.data_ptr() if want_cb else None), N
    tiles = l.fb_gemm_dot_tiles(M, N, K1 + K2, int(bf16), int(force_simt))
    dot = torch.zeros((tiles, M), device=dev) if dotv is not None else None
    g.dotv = dotv.data_ptr() if dotv is not None else None
    g.dot_out = dot.data_ptr() if dot is not None else None
    g.dot_stride = M
    g.M, g.N = M, N
    g.m_dev = m_dev.data_ptr() if m_dev is not None else None
    g.bf16_mode, g.force_simt = int(bf16), int(force_simt)
    _lib.check(l.fb_gemm(C.byref(g), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm")
    torch.cuda.synchronize()
    return Cout, Cb, (dot.sum(0) if dot is not None else None)


def ref_gemm(A, W, bias, act, res, A2, dotv, bf16):
    if bf16:
        A, W = A.to(torch.bfloat16).float(), W.to(torch.bfloat16).float()
        A2 = A2.to(torch.bfloat16).float() if A2 is not None else None
    Af = torch.cat([A, A2], 1) if A2 is not None else A
    v = Af.double() @ W.double().t()
    if bias is not None:
        v = v + bias.double()
    if act == 1:
        v = torch.nn.functional.silu(v)
    elif act == 2:
        v = torch.relu(v)
    if res is not None:
        v = v + res.double()
    d = (v * dotv.double()).sum(1) if dotv is not None else None
    return v, d


CASES = [
    # M, N, K1, K2, act, bias, res, dot
    (37, 64, 32, 0, 0, True, False, False),
    (232, 512, 512, 0, 1, True, False, False),
    (500, 1024, 512, 0, 2, True, False, True),
    (300, 512, 512, 512, 1, True, False, False),
    (129, 96, 48, 48, 0, False, True, False),
    (1000, 32, 512, 0, 0, True, False, False),
    (2824, 512, 512, 0, 1, True, False, True),
    (64, 768, 128, 0, 0, True, True, False),
]


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("case", CASES)
def test_gemm_matches_torch(case, bf16):
    M, N, K1, K2, act, has_b, has_r, has_d = case
    torch.manual_seed(M * 7 + N)
    dev = "cuda"
    A = torch.randn(M, K1, device=dev)
    A2 = torch.randn(M, K2, device=dev) if K2 else None
    W = torch.randn(N, K1 + K2, device=dev) / (K1 + K2) ** 0.5
    bias = torch.randn(N, device=dev) if has_b else None
    res = torch.randn(M, N, device=dev) if has_r else None
    dotv = torch.randn(N, device=dev) if has_d else None
    Cout, Cb, dot = run_gemm(A, W, bias, act, res, A2, dotv, bf16=bf16, want_cb=True)
    ref, dref = ref_gemm(A, W, bias, act, res, A2, dotv, bf16)
    tol = 2e-5
    err = float((Cout.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, f"C rel err {err}"
    cb_tol = 1e-2 if bf16 else tol
    assert float((Cb.double() - ref).abs().max() / ref.abs().max()) < cb_tol
    if has_d:
        derr = float((dot.double() - dref).abs().max() / dref.abs().max())
        assert derr < 5e-5, f"dot rel err {derr}"


def test_gemm_tc_large_and_device_rows():
    """tcgen05 path: many M tiles, device-side row count, row-dot epilogue, against the SIMT kernel and torch"""
    dev = "cuda"
    torch.manual_seed(1)
    M, N, K = 45000, 512, 512
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    dotv = torch.randn(N, device=dev)
    m_dev = torch.tensor([44904], dtype=torch.int32, device=dev)
    C1, Cb1, d1 = run_gemm(A, W, bias, 1, None, None, dotv, m_dev=m_dev, bf16=True, want_cb=True)
    C2, Cb2, d2 = run_gemm(A, W, bias, 1, None, None, dotv, m_dev=m_dev, bf16=True, want_cb=True, force_simt=True)
    ref, dref = ref_gemm(A, W, bias, 1, None, None, dotv, True)
    n = 44904
    assert float((C1[:n].double() - ref[:n]).abs().max() / ref.abs().max()) < 2e-5
    assert float((C1[:n] - C2[:n]).abs().max()) < 1e-4
    assert float((d1[:n].double() - dref[:n]).abs().max() / dref.abs().max()) < 5e-5
    assert torch.isnan(C1[n:]).all()
    assert float((Cb1[:n].float() - ref[:n].float()).abs().max() / ref.abs().max()) < 1e-2


def test_gemm_device_row_count():
    dev = "cuda"
    torch.manual_seed(0)
    A = torch.randn(400, 64, device=dev)
    W = torch.randn(128, 64, device=dev)
    m_dev = torch.tensor([130], dtype=torch.int32, device=dev)
    Cout, _, _ = run_gemm(A, W, m_dev=m_dev)
    ref = A @ W.t()
    assert torch.allclose(Cout[:130], ref[:130], atol=1e-4)
    assert torch.isnan(Cout[130:]).all()   # rows beyond the device-side count are never written


def _params(A, W, bias, act, res, Cout, Cb, M, bf16, dotv=None, dot=None):
    g = _lib.GemmParams()
    K = A.shape[1]
    g.A, g.lda, g.K1 = A.data_ptr(), K, K
    g.A2, g.lda2, g.K2 = None, 0, 0
    g.W = W.data_ptr()
    g.bias = bias.data_ptr() if bias is not None else None
    g.act = act
    g.res, g.ldres = (res.data_ptr() if res is not None else None), W.shape[0]
    g.C, g.ldc = (Cout.data_ptr() if Cout is not None else None), W.shape[0]
    g.Cb, g.ldcb = (Cb.data_ptr() if Cb is not None else None), W.shape[0]
    g.dotv = dotv.data_ptr() if dotv is not None else None
    g.dot_out = dot.data_ptr() if dot is not None else None
    g.dot_stride = M
    g.M, g.N = M, W.shape[0]
    g.m_dev = None
    g.bf16_mode, g.force_simt = int(bf16), 0
    return g


@pytest.mark.parametrize("bf16", [False, True])
@pytest.mark.parametrize("shape", [(232, 2592, 512, 512, 256, 1), (700, 3000, 512, 1024, 1024, 2), (130, 129, 1024, 512, 512, 0),
                                   (90, 500, 64, 96, 32, 1)])
def test_gemm_pair_matches_two_launches(shape, bf16):
    """grouped launch (compound-side rows | protein-side rows of one buffer, different weights) == two layers"""
    M0, M1, K, N0, N1, act = shape
    dev = "cuda"
    torch.manual_seed(M0 + N1)
    dt = torch.bfloat16 if bf16 else torch.float32
    A = torch.randn(M0 + M1, K, device=dev).to(dt).contiguous()
    W0 = (torch.randn(N0, K, device=dev) / K ** 0.5).to(dt).contiguous()
    W1 = (torch.randn(N1, K, device=dev) / K ** 0.5).to(dt).contiguous()
    b0, b1 = torch.randn(N0, device=dev), torch.randn(N1, device=dev)
    # in-place residual on the first problem (C aliases res), typed copy on both
    res0 = torch.randn(M0, N0, device=dev)
    C0 = res0.clone()
    C1 = torch.full((M1, N1), float("nan"), device=dev)
    Cb0 = torch.zeros(M0, N0, dtype=dt, device=dev)
    Cb1 = torch.zeros(M1, N1, dtype=dt, device=dev)
    g0 = _params(A[:M0], W0, b0, act, C0, C0, Cb0, M0, bf16)
    g1 = _params(A[M0:], W1, b1, act, None, C1, Cb1, M1, bf16)
    l = _lib.lib()
    n0 = l.fb_launch_count()
    _lib.check(l.fb_gemm_pair(C.byref(g0), C.byref(g1), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm_pair")
    torch.cuda.synchronize()
    launches = l.fb_launch_count() - n0
    if bf16 and K % 64 == 0 and N0 % 128 == 0 and N1 % 128 == 0:
        assert launches == 1, "the pair was expected to take the grouped tcgen05 launch"
    r0, _ = ref_gemm(A[:M0].float(), W0.float(), b0, act, res0, None, None, bf16)
    r1, _ = ref_gemm(A[M0:].float(), W1.float(), b1, act, None, None, None, bf16)
    for got, ref in ((C0, r0), (C1, r1)):
        assert float((got.double() - ref).abs().max() / ref.abs().max()) < 2e-5
    cb_tol = 1e-2 if bf16 else 2e-5
    for got, ref in ((Cb0, r0), (Cb1, r1)):
        assert float((got.double() - ref).abs().max() / ref.abs().max()) < cb_tol


@pytest.mark.parametrize("M,N,K,act,use_dot,out", [
    (45000, 512, 512, 1, False, "cb"),      # edge MLP second Linear (SiLU, bf16 store)
    (45000, 512, 512, 1, True, "none"),     # coordinate head: row-dot epilogue only
    (44904, 512, 1088, 2, False, "cb"),     # FABind+ edge MLP (K = 17 slabs, ReLU)
    (16384 + 100, 1024, 512, 0, True, "c"),  # last pair tile: the second CTA has no live row; fp32 store
    (99696, 512, 512, 2, True, "cb"),       # FABind+ pair transition (all pair rows), ReLU + row-dot + bf16 store
    (20000, 256, 64, 0, False, "c"),        # one k-slab, one column tile
])
def test_gemm_cta_pair_kernel(M, N, K, act, use_dot, out):
    """the cta_group::2 kernel (gemm_tc4.cu: 256x256 tiles on CTA pairs) on the shapes the stack sends it, against torch"""
    dev = "cuda"
    torch.manual_seed(3)
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    dotv = torch.randn(N, device=dev) if use_dot else None
    C1, Cb1, d1 = run_gemm(A, W, bias, act, None, None, dotv, bf16=True, want_cb=out == "cb", want_c=out == "c")
    ref, dref = ref_gemm(A, W, bias, act, None, None, dotv, True)
    scale = ref.abs().max()
    if out == "c":
        assert float((C1.double() - ref).abs().max() / scale) < 2e-5
    if out == "cb":
        assert float((Cb1.double() - ref).abs().max() / scale) < 1e-2
        assert float((Cb1.double() - ref).abs().mean() / ref.abs().mean()) < 3e-3
    if use_dot:
        assert float((d1.double() - dref).abs().max() / dref.abs().max()) < 5e-5



# ---- split-precision modes (FB_PREC_SPLIT3 / SPLIT6): fp32 operands on the tcgen05 kernels as sums of bf16 planes -------------
def run_gemm_split(A, W, mode, bias=None, act=0, res=None, A2=None, dotv=None, n_split=0):
    l = _lib.lib()
    dev = A.device
    M, K1 = A.shape
    K2 = A2.shape[1] if A2 is not None else 0
    K, N = K1 + K2, W.shape[0]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    Wf = W.contiguous()
    Ws = torch.empty(N, 3 * K, dtype=torch.bfloat16, device=dev)
    _lib.check(l.fb_split_rows(Wf.data_ptr(), K, N, K, Ws.data_ptr(), st), "fb_split_rows")
    planes = Ws.float().view(N, 3, K)
    assert float((planes.sum(1) - Wf).abs().max()) <= 2.0 ** -24 * float(Wf.abs().max())      # w0 + w1 + w2 == w to fp32 rounding
    ws = torch.empty(M * 3 * K, dtype=torch.bfloat16, device=dev)
    g = _lib.GemmParams()
    Ad = A.contiguous()
    A2d = A2.contiguous() if A2 is not None else None
    g.A, g.lda, g.K1 = Ad.data_ptr(), K1, K1
    g.A2, g.lda2, g.K2 = (A2d.data_ptr() if A2d is not None else None), K2, K2
    g.W, g.W_f32 = Ws.data_ptr(), Wf.data_ptr()
    g.split_ws, g.split_ws_bytes = ws.data_ptr(), ws.numel() * 2
    g.bias = bias.data_ptr() if bias is not None else None
    g.act = act
    g.res, g.ldres = (res.data_ptr() if res is not None else None), N
    n_lo = n_split if n_split else N
    Cout = torch.full((M, n_lo), float("nan"), device=dev)
    g.C, g.ldc = Cout.data_ptr(), n_lo
    Chi = torch.full((M, N - n_split), float("nan"), device=dev) if n_split else None
    if n_split:
        g.Cb, g.ldcb, g.n_split = Chi.data_ptr(), N - n_split, n_split
    tiles = l.fb_gemm_dot_tiles(M, N, K, mode, 0)
    dot = torch.zeros((tiles, M), device=dev) if dotv is not None else None
    g.dotv = dotv.data_ptr() if dotv is not None else None
    g.dot_out = dot.data_ptr() if dot is not None else None
    g.dot_stride, g.M, g.N, g.bf16_mode = M, M, N, mode
    _lib.check(l.fb_gemm(C.byref(g), st), "fb_gemm")
    torch.cuda.synchronize()
    out = torch.cat([Cout, Chi], 1) if n_split else Cout
    return out, (dot.sum(0) if dot is not None else None)


SPLIT_CASES = [
    # M, N, K1, K2, act, bias, res, dot, n_split
    (232, 512, 512, 0, 1, True, False, False, 0),        # node-level, SiLU (exact in the split modes)
    (500, 1024, 512, 0, 2, True, False, True, 0),
    (300, 512, 512, 512, 1, True, False, False, 0),      # [h | agg] concatenation
    (777, 512, 128, 0, 0, True, True, False, 0),         # residual
    (20000, 512, 512, 0, 1, True, False, True, 0),       # CTA-pair kernel (M >= 16384, N % 256 == 0), row-dot epilogue
    (3000, 1024 + 128 + 1024, 512, 0, 0, True, False, False, 1024 + 128),   # column-routed outputs (q | k | inter32 || v | vc)
    (900, 1024, 512, 64, 2, True, False, True, 0),       # K2 = 64 (pair transition on [z | t64])
    (37, 64, 32, 0, 0, True, False, False, 0),           # does not tile: FFMA kernel on the fp32 weight
]


@pytest.mark.parametrize("mode", [2, 3])
@pytest.mark.parametrize("case", SPLIT_CASES)
def test_gemm_split_precision(case, mode):
    """bf16x3 (mode 2) keeps terms down to 2^-9 per operand: ~1e-5 of the row scale; six products (mode 3, 'fp32_tc') are fp32-grade:
    <= 2e-6, the bound the FFMA kernel is held to above"""
    M, N, K1, K2, act, has_b, has_r, has_d, n_split = case
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N + mode)
    dev = "cuda"
    A = torch.randn(M, K1, generator=g).to(dev)
    A2 = torch.randn(M, K2, generator=g).to(dev) if K2 else None
    W = (torch.randn(N, K1 + K2, generator=g) / (K1 + K2) ** 0.5).to(dev)
    bias = torch.randn(N, generator=g).to(dev) if has_b else None
    res = torch.randn(M, N, generator=g).to(dev) if has_r else None
    dotv = torch.randn(N, generator=g).to(dev) if has_d else None
    out, dot = run_gemm_split(A, W, mode, bias, act, res, A2, dotv, n_split)
    ref, dref = ref_gemm(A, W, bias, act, res, A2, dotv, False)
    tol = 3e-5 if mode == 2 else 2e-6
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    assert err < tol, (case, mode, err)
    if has_d:
        derr = float((dot.double() - dref).abs().max() / dref.abs().max())
        assert derr < 10 * tol, (case, mode, derr)


def _multi_problem(spec, bf16, dev):
    """spec = (M, N, K1, K2, act, res, out) with out in {"c", "cb", "both", "split"} -> (params, tensors, reference fn)"""
    M, N, K1, K2, act, use_res, out = spec
    dt = torch.bfloat16 if bf16 else torch.float32
    A = torch.randn(M, K1, device=dev).to(dt).contiguous()
    A2 = torch.randn(M, K2, device=dev).to(dt).contiguous() if K2 else None
    W = (torch.randn(N, K1 + K2, device=dev) / (K1 + K2) ** 0.5).to(dt).contiguous()
    b = torch.randn(N, device=dev)
    n_split = (N // 2 + 127) // 128 * 128 if out == "split" else 0        # column routing wants a multiple of 128 (q|k: 1152 of 2176)
    nc = n_split if n_split else N
    res = torch.randn(M, nc, device=dev) if use_res else None
    Cf = res.clone() if use_res else torch.full((M, nc), float("nan"), device=dev)     # in-place residual, like the stack
    Cb = torch.zeros(M, N - n_split, dtype=dt, device=dev)
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K1, K1
    if K2:
        g.A2, g.lda2, g.K2 = A2.data_ptr(), K2, K2
    g.W, g.bias, g.act = W.data_ptr(), b.data_ptr(), act
    if use_res:
        g.res, g.ldres = Cf.data_ptr(), nc
    if out in ("c", "both", "split"):
        g.C, g.ldc = Cf.data_ptr(), nc
    if out in ("cb", "both", "split"):
        g.Cb, g.ldcb = Cb.data_ptr(), N - n_split
    g.M, g.N, g.n_split = M, N, n_split
    g.bf16_mode = int(bf16)
    ref, _ = ref_gemm(A.float(), W.float(), b, act, None, A2.float() if K2 else None, None, bf16)
    if use_res:
        ref = ref + res.double()
    keep = (A, A2, W, b, res, Cf, Cb)
    return g, keep, ref, out, n_split


MULTI_GROUPS = [
    # the folded groups of one layer at the benched shape (forward.cu): rows p = 3216, rows c = 496
    [(3216, 512, 128, 0, 0, True, "both"), (3216, 256, 512, 128, 0, False, "c"), (3216, 1024, 512, 128, 2, False, "cb")],
    [(496, 512, 128, 0, 0, True, "both"), (496, 1024, 512, 128, 2, False, "cb"), (3216, 512, 1024, 0, 0, True, "both")],
    [(496, 512, 1024, 0, 0, True, "both"), (3216, 2176, 512, 0, 0, False, "split"), (496, 2176, 512, 1024, 0, False, "split")],
    [(3712, 512, 512, 0, 0, True, "both"), (496, 512, 512, 512, 0, False, "c"), (3216, 256, 512, 512, 0, False, "c")],
    # ragged: one row, a row count that is not a multiple of the tile, four problems, a single problem
    [(1, 128, 64, 0, 1, False, "cb"), (129, 384, 64, 64, 0, True, "c"), (77, 128, 192, 0, 2, False, "both"), (300, 256, 128, 0, 1, False, "c")],
    [(232, 512, 512, 0, 1, False, "cb")],
]


@pytest.mark.parametrize("prefetch", [0, 1])
@pytest.mark.parametrize("bf16", [True, False])
@pytest.mark.parametrize("group", range(len(MULTI_GROUPS)))
def test_gemm_multi_matches_torch(group, bf16, prefetch):
    """fb_gemm_multi: independent problems with their own operands, K, epilogues and outputs == the layers one by one; in bf16 mode
    the group must take ONE launch (gemm_tc5.cu)"""
    dev = "cuda"
    torch.manual_seed(100 + group)
    probs = [_multi_problem(s, bf16, dev) for s in MULTI_GROUPS[group]]
    arr = (_lib.GemmParams * len(probs))(*[p[0] for p in probs])
    l = _lib.lib()
    n0 = l.fb_launch_count()
    _lib.check(l.fb_gemm_multi(arr, len(probs), prefetch, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm_multi")
    torch.cuda.synchronize()
    if bf16:
        assert l.fb_launch_count() - n0 == 1, "the group was expected to take the multi-problem tcgen05 launch"
    for g, keep, ref, out, n_split in probs:
        Cf, Cb = keep[5], keep[6]
        scale = ref.abs().max()
        if out in ("c", "both"):
            assert float((Cf.double() - ref).abs().max() / scale) < 2e-5
        if out in ("cb", "both"):
            assert float((Cb.double() - ref).abs().max() / scale) < (1e-2 if bf16 else 2e-5)
        if out == "split":
            assert float((Cf.double() - ref[:, :n_split]).abs().max() / scale) < 2e-5
            assert float((Cb.double() - ref[:, n_split:]).abs().max() / scale) < (1e-2 if bf16 else 2e-5)
