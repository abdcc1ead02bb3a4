"""The tcgen05 cross-attention core (csrc/xatt_tc.cu, C ABI fb_row_attention_tc) against the definition of the reference's gated
multi-head attention (FABind/fabind/models/model_utils.py:21-38,96-133 with the gated pair bias of cross_att.py:118-134) evaluated
in torch on the same bf16-rounded projections.  Ragged per-complex blocks, both directions, 1 and 2 key tiles, 1 / 2 / 4 heads per
round.  Tolerance: the probabilities pass through bf16 (2^-9) before P V and the output is stored in bf16."""
import ctypes as C
import math

import pytest
import torch

from fabind_b200 import _lib

pytestmark = pytest.mark.gpu
HD, NH, DH = 128, 4, 32


def _reference(q_is_prot, nc, npr, QG, KV, qcol, gcol, kcol, vcol, PB, Nc_tot):
    """returns O [N, 128] fp32 (rows = internal node id: compound rows first)"""
    N = Nc_tot + sum(npr)
    O = torch.zeros(N, HD)
    c_lo, p_lo, pair0 = 0, 0, 0
    for b in range(len(nc)):
        n_c, n_p = nc[b], npr[b]
        bias = PB[pair0:pair0 + n_c * n_p].view(n_p, n_c, NH)            # [prot, comp, head]
        if q_is_prot:
            q_rows, k_rows, node0 = slice(p_lo, p_lo + n_p), slice(c_lo, c_lo + n_c), Nc_tot + p_lo
            bq = bias.permute(2, 0, 1)
        else:
            q_rows, k_rows, node0 = slice(c_lo, c_lo + n_c), slice(p_lo, p_lo + n_p), c_lo
            bq = bias.permute(2, 1, 0)
        q = QG[q_rows, qcol:qcol + HD].view(-1, NH, DH)
        gt = torch.sigmoid(QG[q_rows, gcol:gcol + HD])
        k = KV[k_rows, kcol:kcol + HD].view(-1, NH, DH)
        v = KV[k_rows, vcol:vcol + HD].view(-1, NH, DH)
        a = torch.softmax(torch.einsum("ihd,jhd->hij", q, k) / math.sqrt(DH) + bq, dim=-1)
        o = torch.einsum("hij,jhd->ihd", a, v).reshape(-1, HD) * gt
        O[node0:node0 + o.shape[0]] = o
        c_lo, p_lo, pair0 = c_lo + n_c, p_lo + n_p, pair0 + n_c * n_p
    return O


@pytest.mark.parametrize("nc,npr", [
    ([31, 31, 31], [201, 201, 201]),            # the benched shape: 32 padded keys (4 heads per round) / 2 key tiles
    ([9, 31, 70, 128], [40, 201, 130, 17]),     # ragged: 1 / 2 / 4 heads per round, two query tiles, a full 128-row compound side
    ([5], [256]),                               # the largest key list the kernel takes
])
def test_attention_core_matches_the_definition(nc, npr):
    l = _lib.lib()
    g = torch.Generator().manual_seed(sum(nc) + sum(npr))
    B, Nc, Np = len(nc), sum(nc), sum(npr)
    CAc = (torch.randn(Nc, 4 * HD, generator=g) * 1.5).to(torch.bfloat16)      # K | V (p-attention)  ||  Q | G (c-attention)
    CAp = (torch.randn(Np, 2 * HD, generator=g) * 1.5).to(torch.bfloat16)      # Q | G (p-attention)
    CAp2 = (torch.randn(Np, 2 * HD, generator=g) * 1.5).to(torch.bfloat16)     # K | V (c-attention)
    P = sum(a * b for a, b in zip(nc, npr))
    PB = torch.randn(2, P, NH, generator=g)
    c_off = torch.tensor([0] + list(torch.tensor(nc).cumsum(0)), dtype=torch.int32)
    p_off = torch.tensor([Nc] + [Nc + int(v) for v in torch.tensor(npr).cumsum(0)], dtype=torch.int32)
    pair_base = torch.tensor([0] + list((torch.tensor(nc) * torch.tensor(npr)).cumsum(0)), dtype=torch.int32)
    dev = "cuda"
    c_off_d, p_off_d, pb_d = c_off.to(dev), p_off.to(dev), pair_base.to(dev)
    CAc_d, CAp_d, CAp2_d, PB_d = CAc.to(dev), CAp.to(dev), CAp2.to(dev), PB.to(dev).contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for q_is_prot in (1, 0):
        O = torch.zeros(Nc + Np, HD, dtype=torch.bfloat16, device=dev)
        if q_is_prot:
            rc = l.fb_row_attention_tc(c_off_d.data_ptr(), p_off_d.data_ptr(), pb_d.data_ptr(), B, Nc, 1, max(npr), max(nc),
                                       CAp_d.data_ptr(), 2 * HD, 0, HD, Np, CAc_d.data_ptr(), 4 * HD, 0, HD, Nc, PB_d[0].data_ptr(),
                                       O.data_ptr(), HD, st)
            ref = _reference(1, nc, npr, CAp.float(), CAc.float(), 0, HD, 0, HD, PB[0], Nc)
            rows = slice(Nc, Nc + Np)
        else:
            rc = l.fb_row_attention_tc(c_off_d.data_ptr(), p_off_d.data_ptr(), pb_d.data_ptr(), B, Nc, 0, max(nc), max(npr),
                                       CAc_d.data_ptr(), 4 * HD, 2 * HD, 3 * HD, Nc, CAp2_d.data_ptr(), 2 * HD, 0, HD, Np,
                                       PB_d[1].data_ptr(), O.data_ptr(), HD, st)
            ref = _reference(0, nc, npr, CAc.float(), CAp2.float(), 2 * HD, 3 * HD, 0, HD, PB[1], Nc)
            rows = slice(0, Nc)
        _lib.check(rc, "fb_row_attention_tc")
        torch.cuda.synchronize()
        got = O.float().cpu()
        err = float((got[rows] - ref[rows]).abs().max() / ref[rows].abs().max())
        assert err < 1.5e-2, (q_is_prot, err)
        other = slice(0, Nc) if q_is_prot else slice(Nc, Nc + Np)
        assert float(got[other].abs().max()) == 0.0          # rows of the key side are not touched


def test_long_key_lists_are_declined():
    l = _lib.lib()
    z = torch.zeros(8, dtype=torch.int32, device="cuda")
    buf = torch.zeros(1024, 512, dtype=torch.bfloat16, device="cuda")
    rc = l.fb_row_attention_tc(z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, 4, 0, 4, 300, buf.data_ptr(), 512, 256, 384, 4, buf.data_ptr(),
                               256, 0, 128, 300, buf.data_ptr(), buf.data_ptr(), 128, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == -4
