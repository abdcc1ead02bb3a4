"""Development probe (needs a -DFB_DIAG build: python -m fabind_b200.build --force --diag): who waits for whom inside the CTA-pair GEMM."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (M, N, K, act, dot, out) in [(44904, 512, 512, 1, False, "bf16"), (44904, 512, 512, 1, True, "none"), (99712, 512, 512, 0, False, "bf16")]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev); dv = torch.randn(N, device=dev)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    nt = l.fb_gemm_dot_tiles(M, N, K, 1, 0)
    dout = torch.zeros(nt * M, device=dev)
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = act
    if out == "bf16":
        g.Cb, g.ldcb = Cb.data_ptr(), N
    if dot:
        g.dotv = dv.data_ptr(); g.dot_out = dout.data_ptr(); g.dot_stride = M
    g.M, g.N = M, N; g.bf16_mode = 1
    for _ in range(3):
        l.fb_gemm(C.byref(g), st)
    dbg = torch.zeros(8192, dtype=torch.int64, device=dev)
    l.fb_gemm_set_debug(C.c_void_p(dbg.data_ptr())); l.fb_gemm(C.byref(g), st); torch.cuda.synchronize(); l.fb_gemm_set_debug(None)
    d = dbg[:148 * 8].view(148, 8).cpu().double()
    lead, peer = d[0::2], d[1::2]
    names = ["mma_wait_full", "mma_wait_tempty", "prod0_wait_empty", "epi0_wait_tfull", "epi0_busy", "kernel_total"]
    print(json.dumps(dict(M=M, dot=dot, out=out, leader_mean_clk={n: round(lead[:, i].mean().item()) for i, n in enumerate(names)},
                          peer_mean_clk={n: round(peer[:, i].mean().item()) for i, n in enumerate(names)})))
