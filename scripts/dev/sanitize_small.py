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
# post-optimisation
ref = torch.randn(40, 3, device="cuda"); pred = ref + 0.3 * torch.randn(40, 3, device="cuda")
batch = torch.cat([torch.zeros(15), torch.ones(25)]).long().cuda()
las = torch.tensor([[0, 1, 2, 0, 3], [1, 0, 3, 2, 4]]).cuda(); lb = torch.tensor([0, 0, 0, 1, 1]).cuda()
post_optimize_batch(ref, pred, batch, las, lb, total_epoch=20)
torch.cuda.synchronize()
print("sanitize_small: done")
