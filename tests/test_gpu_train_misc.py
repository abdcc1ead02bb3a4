"""Training-step helpers on the GPU: the batched transposed bf16 twins of the weight arena (fb_transpose_slots_bf16) against torch."""
import pytest
import torch

from fabind_b200 import train, backward as bw
from fabind_b200.weights import slots, base_elems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("hidden,layers,flavour", [(128, 2, 0), (512, 4, 0), (64, 2, 1)])
def test_transposed_twins_match_torch(hidden, layers, flavour):
    from fabind_b200 import _lib
    n = _lib.lib().fb_weight_arena_elems_f(hidden, layers, flavour)
    g = torch.Generator().manual_seed(3)
    arena = torch.randn(n, generator=g).cuda()
    old = bw.PRECISION
    bw.PRECISION = "bf16"
    try:
        w = train.slot_tensors(arena, hidden, layers, flavour)
    finally:
        bw.PRECISION = old
    torch.cuda.synchronize()
    seen = 0
    for name, r, c, off in slots(hidden, layers, flavour):
        if r <= 1 or r * c == 0:
            continue
        pre, _, base = name.rpartition(".")
        pre = pre + "." if pre else ""
        t = w[pre][base + "_t"]
        want = arena[off:off + r * c].view(r, c).to(torch.bfloat16).t().contiguous()
        assert t.shape == (c, r) and t.dtype == torch.bfloat16 and t.is_contiguous()
        assert torch.equal(t, want), name
        seen += 1
    assert seen > 10 and base_elems(hidden, layers, flavour) <= n


def test_row_split_weight_gradient_matches_torch():
    """backward.gemm_wgrad on tcgen05: the reduction over the rows split into independent problems of one multi-problem launch
    (strided W operand, fb_gemm_params.ldw) against dY^T X in fp32 on bf16-rounded operands"""
    old = bw.PRECISION
    bw.PRECISION = "bf16"
    try:
        g = torch.Generator().manual_seed(5)
        for M, N, K in [(44904, 512, 512), (9000, 128, 512), (4100, 512, 640), (3000, 512, 512)]:
            dY = torch.randn(M, N, generator=g).cuda()
            X = torch.randn(M, K, generator=g).cuda()
            got = bw.gemm_wgrad(dY, X)
            want = dY.to(torch.bfloat16).float().t() @ X.to(torch.bfloat16).float()
            torch.cuda.synchronize()
            assert got.shape == (N, K)
            assert float((got - want).abs().max()) <= 2e-3 * float(want.abs().max()), (M, N, K)
            acc = bw.gemm_wgrad(dY, X, out=got.clone())
            assert float((acc - 2 * want).abs().max()) <= 4e-3 * float(want.abs().max())
    finally:
        bw.PRECISION = old
