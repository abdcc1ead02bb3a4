"""Development probe: device time of the long GEMM shapes with the CTA-pair kernel (default) or the v3 kernel (FB_TC4=0)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for (M, N, K, act, dot) in [(44922, 512, 512, 1, False), (44922, 512, 512, 1, True), (99696, 512, 512, 2, True), (44922, 512, 1088, 2, False)]:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev); dv = torch.randn(N, device=dev)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = act
    if dot:
        tiles = l.fb_gemm_dot_tiles(M, N, K, 1, 0)
        d = torch.zeros(tiles, M, device=dev)
        g.dotv, g.dot_out, g.dot_stride = dv.data_ptr(), d.data_ptr(), M
    else:
        g.Cb, g.ldcb = Cb.data_ptr(), N
    g.M, g.N = M, N; g.bf16_mode = 1
    for _ in range(3):
        l.fb_gemm(C.byref(g), st)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); l.fb_gemm(C.byref(g), st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    us = ts[len(ts) // 2] * 1e3
    out[f"{M}x{N}x{K}{'+dot' if dot else ''}"] = dict(us=round(us, 1), tflops=round(2.0 * M * N * K / us / 1e6, 1))
print(json.dumps(dict(tc4=os.environ.get("FB_TC4", "1"), **out)))
