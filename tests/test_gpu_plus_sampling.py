"""FABind+ sampling mode on the GPU: train() under no_grad runs the stack with every nn.Dropout site active (the reference's
`--infer-dropout` path, P/test_sampling_fabind.py:118-124).  The masks are the library's counter-based function
(fabind_b200/dropout.py); parity is pinned two ways:
 (a) column-only masks against the UNMODIFIED reference in train() mode with its nn.Dropout modules patched to the same masks
     (placement and scaling of all 17 sites per layer), and
 (b) full row x column masks against the CPU emulation of the launch sequence with the same mask function (row identities)."""
import glob
import os

import pytest
import torch

from oracle import ref_shims
from fabind_b200.plus import EfficientMCAttModel
from helpers import GOLDEN_DIR, load_golden, rel_err
from emulate_packed import forward_emulated

pytestmark = pytest.mark.gpu
FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "plusdrop_*.pt")))


def _model(r, sd, precision="fp32"):
    args = ref_shims.published_args_plus(dropout=r["dropout_p"], random_n_iter=False)
    m = EfficientMCAttModel(args, r["hidden"], r["hidden"], 1, n_layers=r["n_layers"], n_iter=r["n_iter"],
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    m.precision = precision
    m.dropout_seed = r["dropout_seed"]
    return m


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-3])
def test_column_masks_match_patched_reference(path):
    g, r, b, sd, cfg = load_golden(path)
    m = _model(r, sd)
    m.dropout_colonly = True
    with torch.no_grad():
        X, H, pair = m(**b.to("cuda").forward_args())
    assert rel_err(X, g["X"]) < 1e-4 and rel_err(H, g["H"]) < 1e-4 and rel_err(pair, g["pair"]) < 1e-4


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("/")[-1][:-3])
def test_full_masks_match_emulation(path):
    g, r, b, sd, cfg = load_golden(path)
    with torch.no_grad():
        Xe, He, _, _ = forward_emulated(sd, cfg, b, flavour=1, dropout=(r["dropout_p"], r["dropout_seed"], False))
    m = _model(r, sd)
    with torch.no_grad():
        X, H, pair = m(**b.to("cuda").forward_args())
    assert rel_err(X, Xe) < 1e-4 and rel_err(H, He) < 1e-4
    assert rel_err(H, g["H"]) > 1e-2       # different masks than the column-only golden: really a different sample


def test_sampling_mode_semantics():
    g, r, b, sd, cfg = load_golden(FILES[0])
    m = _model(r, sd, precision="bf16")
    with pytest.raises(NotImplementedError):      # train() with autograd on = training, which is not built
        m(**b.to("cuda").forward_args())
    outs = []
    for seed in (1, 1, 2):
        m.dropout_seed = seed
        with torch.no_grad():
            X, H, _ = m(**b.to("cuda").forward_args())
        outs.append((X.clone(), H.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])      # same seed: same sample, bit for bit
    assert float((outs[0][1] - outs[2][1]).abs().max()) > 1e-3                               # another seed: another sample
    m.dropout_seed = None                                                                     # unseeded: torch's generator
    torch.manual_seed(0)
    with torch.no_grad():
        a = m(**b.to("cuda").forward_args())[1].clone()
        c = m(**b.to("cuda").forward_args())[1].clone()
    assert float((a - c).abs().max()) > 1e-3
    m.eval()
    with torch.no_grad():
        e1 = m(**b.to("cuda").forward_args())[1].clone()
        e2 = m(**b.to("cuda").forward_args())[1].clone()
    assert torch.equal(e1, e2)


def test_gemm_dropout_keep_rate():
    """fb_gemm with the dropout epilogue: kept fraction ~ 1 - p, kept values scaled by 1/(1-p), every GEMM kernel agrees"""
    import ctypes as C
    from fabind_b200 import _lib
    l = _lib.lib()
    res = {}
    for name, (M, bf, simt) in dict(simt=(3000, False, True), tc3=(3000, True, False), tc4=(20000, True, False)).items():
        torch.manual_seed(0)
        A = torch.randn(M, 512, device="cuda")
        W = torch.randn(512, 512, device="cuda") / 512 ** 0.5
        dt = torch.bfloat16 if bf else torch.float32
        Ad, Wd = A.to(dt).contiguous(), W.to(dt).contiguous()
        out = {}
        for p in (0.0, 0.25):
            Cout = torch.zeros(M, 512, device="cuda")
            q = _lib.GemmParams()
            q.A, q.lda, q.K1 = Ad.data_ptr(), 512, 512
            q.W, q.act = Wd.data_ptr(), 0
            q.C, q.ldc, q.M, q.N = Cout.data_ptr(), 512, M, 512
            q.bf16_mode, q.force_simt = int(bf), int(simt)
            q.drop_p, q.drop_seed, q.drop_site, q.drop_row0, q.drop_colonly = p, 77, 5, 11, 0
            _lib.check(l.fb_gemm(C.byref(q), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm")
            torch.cuda.synchronize()
            out[p] = Cout
        kept = out[0.25] != 0
        assert abs(float(kept.float().mean()) - 0.75) < 0.01, name
        assert torch.allclose(out[0.25][kept], out[0.0][kept] / 0.75, rtol=1e-5, atol=1e-6), name
        res[name] = kept[:3000].cpu()
    assert torch.equal(res["simt"], res["tc3"]) and torch.equal(res["simt"], res["tc4"])       # one mask function everywhere
    from fabind_b200.dropout import keep_mask
    assert torch.equal(res["simt"], keep_mask(77, 5, 3000, 512, 0.25, row0=11) > 0)


def test_wrapper_sampling_matches_patched_reference():
    """FABindPlus.inference in sampling mode (train(), ranking modules in eval, DBSCAN pocket clustering, confidence head,
    random_n_iter draws from python `random`): coordinates and confidence scores against the UNMODIFIED reference run the same
    way with its nn.Dropout modules patched to the library's column-only masks (scripts/make_golden.py::main_l2_plus_sampling)."""
    import random
    from fabind_b200.config import published_args_plus
    from fabind_b200.plus import FABindPlus
    from fabind_b200.synthetic import make_docking_batch
    from oracle.det_weights import det_state_dict
    g = torch.load(os.path.join(GOLDEN_DIR, "l2plussample_h64_p32_l2_it2.pt"), map_location="cpu", weights_only=False)
    r = g["recipe"]
    args = published_args_plus(mean_layers=r["mean_layers"], n_iter=r["n_iter"], dropout=r["dropout_p"], confidence_training=True,
                               stack_mlp=True, use_clustering=True, random_n_iter=True)
    m = FABindPlus(args, r["emb"], r["pemb"])
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == dict(g["shapes"])
    m.load_state_dict(det_state_dict(g["shapes"], r["weight_seed"]), strict=True)
    m = m.cuda().train()
    for name, sub in m.named_modules():
        if name.startswith("confidence") or name.startswith("ranking"):
            sub.eval()
    m.dropout_seed, m.dropout_colonly = r["dropout_seed"], True
    data = make_docking_batch(**r["batch"]).to("cuda")
    random.seed(r["random_seed"])
    with torch.no_grad():
        coords, batch, conf = m.inference(data)
    assert rel_err(coords, g["coords"]) < 1e-4
    assert rel_err(conf, g["confidence"]) < 1e-4
    # sample(): n passes with different masks -> different poses, the API the sampling scripts need
    outs = m.sample(lambda: make_docking_batch(**r["batch"]).to("cuda"), 3, seed=5)
    assert len(outs) == 3 and float((outs[0][0] - outs[1][0]).abs().max()) > 1e-4


def test_batched_sampling_equals_replicated_rows():
    """sample_batched: S samples = S replicas in one batch.  Replica k of a complex must equal what a stand-alone pass over the
    replicated batch gives (trivially) AND be a valid independent sample: with dropout off (p = 0) all replicas coincide with
    the eval-mode pose; with dropout on they differ from each other."""
    from fabind_b200.config import published_args_plus
    from fabind_b200.plus import FABindPlus
    from fabind_b200.plus.sampling import replicate_batch, sample_batched
    from fabind_b200.synthetic import make_docking_batch
    from oracle.det_weights import det_state_dict
    args = published_args_plus(mean_layers=1, n_iter=2, confidence_training=True, stack_mlp=True, random_n_iter=False)
    m = FABindPlus(args, 64, 32)
    m.load_state_dict(det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 71), strict=True)
    m = m.cuda().eval()
    data = make_docking_batch(2, seed=21, n_c_range=(8, 16), L_range=(100, 160)).to("cuda")
    with torch.no_grad():
        ref_coords, _, ref_conf = m.inference(data)
        rep = m.inference(replicate_batch(data, 3))                   # eval mode: replicas are exact copies
    n = ref_coords.shape[0]
    for k in range(3):
        assert rel_err(rep[0][k * n:(k + 1) * n], ref_coords) < 1e-5 and rel_err(rep[2][k * 2:(k + 1) * 2], ref_conf) < 1e-5
    coords, batch, conf = sample_batched(m, data, 5, seed=3, max_instances=6)     # chunks of 3 + 2 samples
    assert coords.shape == (5, n, 3) and conf.shape == (5, 2) and not m.training
    d01 = float((coords[0] - coords[1]).abs().max())
    assert d01 > 1e-4 and float((coords[0] - ref_coords).abs().max()) > 1e-4


def test_wrapper_train_mode_forward_matches_patched_reference():
    """FABindPlus.forward(data, stage=2) in train() mode (the call of test_sampling_fabind.py's validate()): gumbel-softmax pocket
    centre on injected noise + column-only dropout masks, all 13 outputs against the unmodified reference run the same way."""
    from fabind_b200.config import published_args_plus
    from fabind_b200.plus import FABindPlus
    from fabind_b200.synthetic import make_docking_batch
    from oracle.det_weights import det_state_dict
    g = torch.load(os.path.join(GOLDEN_DIR, "l2plustrainfwd_h64_p32_l2_it2.pt"), map_location="cpu", weights_only=False)
    r = g["recipe"]
    args = published_args_plus(mean_layers=r["mean_layers"], n_iter=r["n_iter"], dropout=r["dropout_p"], random_n_iter=False)
    m = FABindPlus(args, r["emb"], r["pemb"])
    m.load_state_dict(det_state_dict(g["shapes"], r["weight_seed"]), strict=True)
    m = m.cuda().train()
    m.dropout_seed, m.dropout_colonly, m.gumbel_noise = r["dropout_seed"], True, g["noise"]
    data = make_docking_batch(**r["batch"]).to("cuda")
    with torch.no_grad():
        out = m(data, stage=2)
    ref = g["forward"]
    assert len(out) == len(ref) == 13
    for i, (a, b) in enumerate(zip(out, ref)):
        if torch.is_tensor(b):
            assert tuple(a.shape) == tuple(b.shape), i
            if b.dtype.is_floating_point:
                if i == 11:
                    assert float((a.cpu() - b).abs().max()) < 1e-3, i       # relu(radius head) near its threshold
                else:
                    assert rel_err(a, b) < 1e-4, (i, rel_err(a, b))
            else:
                assert torch.equal(a.cpu().to(b.dtype), b), i
        else:
            assert a == b, i
    assert rel_err(data.coords.cpu(), g["coords_after"]) < 1e-5
