"""compute-sanitizer target: one small forward of every product path (v1 / FABind+ stacks in both precisions incl. sampling mode,
L2 wrappers, CTA-pair GEMM, post-optimisation).  Run:  compute-sanitizer --tool memcheck python scripts/dev/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import EfficientMCAttModel, _lib
from fabind_b200.config import published_args, published_args_plus
from fabind_b200.model import IaBNet_mean_and_pocket_prediction_cls_coords_dependent as Net
from fabind_b200.plus import EfficientMCAttModel as PlusModel, FABindPlus
from fabind_b200.post_optim import post_optimize_batch
from fabind_b200.synthetic import make_batch, make_docking_batch, randomize_coord_heads

torch.manual_seed(0)
nc = dict(normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
b = make_batch(n_complexes=3, seed=1, embed=128, n_c_range=(6, 20), n_p_range=(30, 70))
for cls, args in ((EfficientMCAttModel, published_args()), (PlusModel, published_args_plus(random_n_iter=False))):
    m = cls(args, 128, 128, 1, n_layers=2, n_iter=2, **nc)
    randomize_coord_heads(m)
    m = m.cuda().eval()
    for prec in ("fp32", "bf16"):
        m.precision = prec
        m(**b.to("cuda").forward_args())
    if cls is PlusModel:
        m.train(); m.dropout_seed = 3
        with torch.no_grad():
            m(**b.to("cuda").forward_args())
d = make_docking_batch(2, seed=2, n_c_range=(6, 14), L_range=(60, 120))
n1 = Net(published_args(mean_layers=1, n_iter=2), 128, 64).cuda().eval()
n1.precision = "bf16"; n1(d.to("cuda"), stage=2); n1.inference(d.to("cuda"))
n2 = FABindPlus(published_args_plus(mean_layers=1, n_iter=2, confidence_training=True, stack_mlp=True, use_clustering=True, random_n_iter=False), 128, 64).cuda().eval()
n2.precision = "bf16"; n2(d.to("cuda"), stage=2); n2.inference(d.to("cuda"))
# CTA-pair GEMM
M, N, K = 16500, 256, 128
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
Cb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
q = _lib.GemmParams()
q.A, q.lda, q.K1, q.W, q.Cb, q.ldcb, q.M, q.N, q.bf16_mode, q.act = A.data_ptr(), K, K, W.data_ptr(), Cb.data_ptr(), N, M, N, 1, 1
_lib.check(_lib.lib().fb_gemm(C.byref(q), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm")
# round 2: the streaming form of the CTA-pair GEMM (K > 512), row-dot epilogue, fp32 output
M, N, K = 16600, 256, 576
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
Cf = torch.empty(M, N, device="cuda"); dv = torch.randn(N, device="cuda")
nt = _lib.lib().fb_gemm_dot_tiles(M, N, K, 1, 0)
dout = torch.zeros(nt * M, device="cuda")
q = _lib.GemmParams()
q.A, q.lda, q.K1, q.W, q.C, q.ldc, q.M, q.N, q.bf16_mode, q.act = A.data_ptr(), K, K, W.data_ptr(), Cf.data_ptr(), N, M, N, 1, 2
q.dotv, q.dot_out, q.dot_stride = dv.data_ptr(), dout.data_ptr(), M
_lib.check(_lib.lib().fb_gemm(C.byref(q), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm")
# round 2: multi-problem launch (gemm_tc5.cu): K-concatenated operands, residual, column-routed outputs, ragged row counts
def prob(M, N, K1, K2, act, res, split):
    g = _lib.GemmParams()
    A = torch.randn(M, K1, device="cuda").to(torch.bfloat16); A2 = torch.randn(M, max(K2, 8), device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K1 + K2, device="cuda").to(torch.bfloat16); b = torch.randn(N, device="cuda")
    ns = 128 if split else 0
    Cf = torch.randn(M, ns if ns else N, device="cuda"); Cb = torch.zeros(M, N - ns, device="cuda", dtype=torch.bfloat16)
    g.A, g.lda, g.K1, g.W, g.bias, g.act = A.data_ptr(), K1, K1, W.data_ptr(), b.data_ptr(), act
    if K2:
        g.A2, g.lda2, g.K2 = A2.data_ptr(), K2, K2
    g.C, g.ldc, g.Cb, g.ldcb = Cf.data_ptr(), Cf.shape[1], Cb.data_ptr(), Cb.shape[1]
    if res:
        g.res, g.ldres = Cf.data_ptr(), Cf.shape[1]
    g.M, g.N, g.n_split, g.bf16_mode = M, N, ns, 1
    return g, (A, A2, W, b, Cf, Cb)
ps = [prob(300, 256, 128, 64, 2, False, True), prob(129, 128, 64, 0, 0, True, False), prob(1, 384, 192, 0, 1, False, False), prob(700, 128, 128, 128, 0, True, False)]
arr = (_lib.GemmParams * 4)(*[p[0] for p in ps])
for pre in (0, 1):
    _lib.check(_lib.lib().fb_gemm_multi(arr, 4, pre, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb_gemm_multi")
# round 2: folded sequences + iteration-invariant head + moving-rows out layer (three iterations), tcgen05 attention core
m = EfficientMCAttModel(published_args(), 128, 128, 1, n_layers=2, n_iter=3, **nc)
randomize_coord_heads(m)
m = m.cuda().eval()
for prec, att in (("bf16", "simt"), ("fp32", "simt"), ("fp32_tc", "simt"), ("bf16", "tcgen05")):
    m.precision, m.attention = prec, att
    m(**b.to("cuda").forward_args())
# round 2 (late): dataloader-side layout -- a correct hint, then WRONG hints (too few / too many context edges: the device clamps its row
# pointers into the claimed sizes and zero-fills the edge lists; flagged garbage, never an out-of-bounds access)
from fabind_b200 import runtime
from fabind_b200.dataloader import layout_hint, attach, prepare_batch
m.precision, m.attention = "bf16", "simt"
m(**prepare_batch(b.forward_args(), "cuda", m.layout_cutoff()))
for de, dm in ((-41, 0), (+57, 0), (0, -7), (+9, +9)):
    h = layout_hint(b.X, b.batch_id, b.segment_id, b.mask, b.is_global, b.compound_edge_index, m.layout_cutoff())
    h.e_ctx += de; h.e_ctx_mv += dm
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in b.forward_args().items()}
    attach(h, dev["batch_id"], dev["segment_id"], dev["is_global"], dev["mask"], "cuda")
    m(**dev)
    torch.cuda.synchronize()
    runtime.layout_flag("cuda").zero_()
# a batch at the benched per-complex sizes: head CTAs of gcl_node (global rows with 31 / 201 edges), four-query row attention
b2 = make_batch(n_complexes=16, n_c=30, n_p=200, embed=128, seed=5)
m(**b2.to("cuda").forward_args())
# training step through the drop-in module (fused training-forward kernels, batched slot transposes, packer)
from fabind_b200 import backward as bw
mt = EfficientMCAttModel(published_args(), 128, 128, 1, n_layers=2, n_iter=2, **nc)
randomize_coord_heads(mt)
mt = mt.cuda().train(); mt.precision = "bf16"; mt.dropout_seed = 7
bw.PRECISION = "bf16"
for _ in range(2):
    X, H = mt(**b.to("cuda").forward_args())
    (X.sum() + H.sum()).backward()
bw.PRECISION = "fp32"
# post-optimisation
ref = torch.randn(40, 3, device="cuda"); pred = ref + 0.3 * torch.randn(40, 3, device="cuda")
batch = torch.cat([torch.zeros(15), torch.ones(25)]).long().cuda()
las = torch.tensor([[0, 1, 2, 0, 3], [1, 0, 3, 2, 4]]).cuda(); lb = torch.tensor([0, 0, 0, 1, 1]).cuda()
post_optimize_batch(ref, pred, batch, las, lb, total_epoch=20)
torch.cuda.synchronize()
print("sanitize_small: done")
