"""Dataloader-side layout on the GPU (fb_model_params.layout_flag, ABI 6): a forward whose edge counts came from the collate step
gives the same bits as the read-back protocol, leaves the flag clear and performs no device->host transfer before its outputs; a
wrong claim is flagged (and raised in host-buffer mode) without touching memory outside the buffers the host sized."""
import pytest
import torch

from oracle import ref_shims
from oracle.det_weights import det_state_dict
from fabind_b200 import EfficientMCAttModel, layout, runtime
from fabind_b200.dataloader import layout_hint, attach, prepare_batch
from fabind_b200.synthetic import make_batch

pytestmark = pytest.mark.gpu


def _model(hidden=128, L=2, IT=3, precision="fp32"):
    m = EfficientMCAttModel(ref_shims.published_args(), hidden, hidden, 1, n_layers=L, n_iter=IT,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    m.precision = precision
    return m


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_hinted_forward_is_bit_identical_and_sync_free(precision):
    m = _model(precision=precision)
    b = make_batch(n_complexes=4, seed=11, embed=128, n_c_range=(8, 40), n_p_range=(40, 150))
    plain = b.to("cuda")
    X0, H0 = m(**plain.forward_args())
    e0 = m.last_stats["ctx_edges"]
    args = prepare_batch(b.forward_args(), "cuda", m.layout_cutoff())
    lay = layout.build_layout(args["batch_id"], args["segment_id"], args["is_global"], args["mask"], "cuda")
    assert lay.e_ctx == e0
    torch.cuda.synchronize()
    m(**{k: (v.clone() if k == "X" else v) for k, v in args.items()})     # warm: weight arenas, scratch
    torch.cuda.synchronize()
    xin = args["X"].clone()
    torch.cuda.set_sync_debug_mode("error")       # any implicit synchronisation from torch raises
    try:
        X1, H1 = m(**dict(args, X=xin))
    finally:
        torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
    assert torch.equal(X1, X0) and torch.equal(H1, H0)
    runtime.check_layout_flag("cuda")             # clear: the claim held


def test_wrong_hint_is_flagged_not_fatal():
    m = _model()
    b = make_batch(n_complexes=3, seed=12, embed=128, n_c_range=(8, 30), n_p_range=(40, 120))
    good = layout_hint(b.X, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, m.layout_cutoff())
    X_ref, H_ref = m(**b.to("cuda").forward_args())
    for de, dm in ((-37, 0), (+64, 0), (0, -5), (+8, +8)):
        bad = layout_hint(b.X, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, m.layout_cutoff())
        bad.e_ctx += de
        bad.e_ctx_mv += dm
        host = b.forward_args()                                    # host-buffer mode: checked at the final synchronisation
        host = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in host.items()}
        attach(bad, host["batch_id"], host["segment_id"], host["is_global"], host["mask"], "cuda")
        with pytest.raises(RuntimeError, match="dataloader-side layout"):
            m(**host)
        runtime.check_layout_flag("cuda")                          # raised once, cleared
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.forward_args().items()}
        attach(bad, dev["batch_id"], dev["segment_id"], dev["is_global"], dev["mask"], "cuda")
        m(**dev)                                                   # device-resident: never synchronises, flag only
        with pytest.raises(RuntimeError):
            runtime.check_layout_flag("cuda")
    # and the library is intact afterwards
    host = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b.forward_args().items()}
    attach(good, host["batch_id"], host["segment_id"], host["is_global"], host["mask"], "cuda")
    X2, H2 = m(**host)
    assert torch.equal(X2, X_ref.cpu()) and torch.equal(H2, H_ref.cpu())
