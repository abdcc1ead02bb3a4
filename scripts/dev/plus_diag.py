"""Development aid (FABind+ layout): (1) per-sub-layer deviation of bf16 mode from fp32 mode on the GPU, (2) per-stage
CUDA-event breakdown of one forward at the config-4 per-GPU shape."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import ref_shims
from oracle.det_weights import det_state_dict
from fabind_b200 import _lib
from fabind_b200.plus import EfficientMCAttModel
from fabind_b200.synthetic import make_batch, randomize_coord_heads

CATS = ["gemm_edge", "gemm_node", "gemm_pair", "gemm_pair0", "edge_elementwise", "attention", "graph_misc"]


def build(L, IT, det):
    m = EfficientMCAttModel(ref_shims.published_args_plus(), 512, 512, 1, n_layers=L, n_iter=IT,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    if det:
        m.load_state_dict(det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 33), strict=True)
    else:
        torch.manual_seed(0)
        randomize_coord_heads(m)
    return m.cuda().eval()


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


if "diag" in sys.argv:
    for det in (True, False):
        L, IT = 5, 1
        m = build(L, IT, det)
        m.debug_trace = True
        out = {}
        for prec in ("fp32", "bf16"):
            b = make_batch(embed=512, n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200)).to("cuda")
            m.precision = prec
            X, H, pair = m(**b.forward_args())
            th, tx = m.last_stats["trace"]
            out[prec] = (X.clone(), H.clone(), pair.clone(), th.clone(), tx.clone())
        a, bq = out["fp32"], out["bf16"]
        rec = dict(det_weights=det, X=rel(bq[0], a[0]), H=rel(bq[1], a[1]), pair=rel(bq[2], a[2]),
                   h_mean_over_std=float(a[1].mean(1).abs().mean() / a[1].std(1).mean()),
                   pair_mean_over_std=float(a[2][0].mean(-1).abs().mean() / a[2][0].std(-1).mean()))
        for k in range(2 * L):
            rec[f"{'gcl' if k % 2 == 0 else 'att'}_{k // 2}"] = (round(rel(bq[3][k], a[3][k]), 5), round(rel(bq[4][k], a[4][k]), 5),
                                                                  round(float(a[3][k].mean(1).abs().mean() / a[3][k].std(1).mean()), 3))
        print(json.dumps(rec))

if "stages" in sys.argv:
    lib = _lib.lib()
    L, IT, B = 5, 8, 16
    m = build(L, IT, False)
    m.precision = "bf16"
    m.return_pair = False
    b = make_batch(n_complexes=B, seed=0, n_c=30, n_p=200).to("cuda")
    X0 = b.X.clone()
    for _ in range(2):
        b.X.copy_(X0); m(**b.forward_args())
    torch.cuda.synchronize()
    lib.fb_prof_enable(1)
    n = 3
    for _ in range(n):
        b.X.copy_(X0); m(**b.forward_args())
    torch.cuda.synchronize()
    ms = (C.c_double * len(CATS))(); spans = (C.c_int64 * len(CATS))()
    lib.fb_prof_read(ms, spans, len(CATS))
    lib.fb_prof_enable(0)
    print(json.dumps({c: (round(ms[i] / n, 3), spans[i] // n) for i, c in enumerate(CATS)}))

if "iters" in sys.argv:
    # tiny (reference-initialised) coordinate heads: H of the last iteration must deviate the same for any n_iter
    for IT in (1, 2, 3):
        m = build(5, IT, True)
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        g = torch.Generator().manual_seed(7)
        for k in sd:
            if k.endswith("coord_mlp.linear2.weight"):
                sd[k] = ((torch.rand(sd[k].shape, generator=g) * 2 - 1) * 0.001 * (6.0 / 513) ** 0.5).to(sd[k].device)
        m.load_state_dict(sd)
        out = {}
        for prec in ("fp32", "bf16"):
            b = make_batch(embed=512, n_complexes=2, seed=5, n_c_range=(10, 50), n_p_range=(80, 200)).to("cuda")
            m.precision = prec
            X, H, pair = m(**b.forward_args())
            out[prec] = (X.clone(), H.clone(), pair.clone())
        a, bq = out["fp32"], out["bf16"]
        d = (bq[1] - a[1]).abs()
        print(json.dumps(dict(IT=IT, X=rel(bq[0], a[0]), H=rel(bq[1], a[1]), pair=rel(bq[2], a[2]), H_mean_abs_err=float(d.mean()),
                              H_scale=float(a[1].abs().max()), H_rms=float(a[1].pow(2).mean().sqrt()),
                              worst_row=int(d.max(1).values.argmax()), n_bad=int((d.max(1).values > 0.01 * a[1].abs().max()).sum()))))
