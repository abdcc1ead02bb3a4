"""Development probe: phase timeline of the persistent tcgen05 GEMM (first tile of sampled CTAs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K, act, both) in [(3712, 512, 512, 0, False), (3712, 512, 512, 0, True), (496, 512, 512, 0, False), (3712, 1024, 512, 2, False), (44922, 512, 512, 1, False)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev); Cf = torch.empty(M, N, device=dev)
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = act
    g.Cb, g.ldcb = Cb.data_ptr(), N
    if both:
        g.C, g.ldc = Cf.data_ptr(), N; g.res, g.ldres = Cf.data_ptr(), N
    g.M, g.N = M, N; g.bf16_mode = 1
    for _ in range(3):
        l.fb_gemm(C.byref(g), st)
    dbg = torch.zeros(8192, dtype=torch.int64, device=dev)
    l.fb_gemm_set_debug(C.c_void_p(dbg.data_ptr())); l.fb_gemm(C.byref(g), st); torch.cuda.synchronize(); l.fb_gemm_set_debug(None)
    d = dbg[:148 * 8].view(-1, 8).cpu().double(); d = d[d[:, 0] > 0]
    rel = d[:, 1:7] - d[:, 0:1]
    names = ["setup done", "first TMA landed", "tile0 MMA issued", "tile0 accum ready", "tile0 epilogue done", "exit"]
    print(json.dumps(dict(M=M, N=N, K=K, act=act, both=both, ctas=int(d.shape[0]),
                          mean_ns={n: round(rel[:, i].mean().item()) for i, n in enumerate(names)})))
