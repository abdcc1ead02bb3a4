"""Development probe: event-timed tcgen05 GEMM for every shape the docking stack launches (B=16)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib

l = _lib.lib()
dev = "cuda"
SHAPES = [  # name, M, N, K, act, dot, m_dev
    ("edge e2", 44922, 512, 512, 1, False, None), ("edge c1+dot", 44922, 512, 512, 1, True, None),
    ("pair0", 100496, 512, 512, 0, False, None), ("pair pt1+dot", 96000, 1024, 512, 2, True, 11200),
    ("node Pn", 3712, 1024, 512, 0, False, None), ("node n1 K=1024", 3712, 512, 1024, 1, False, None),
    ("node n2", 3712, 512, 512, 0, False, None), ("node qk", 3712, 1152, 512, 0, False, None),
    ("ca_c", 496, 512, 512, 0, False, None), ("ca_p", 3216, 256, 512, 0, False, None),
    ("o_p K=128", 3216, 512, 128, 0, False, None), ("tp1", 3216, 1024, 512, 2, False, None),
    ("tp2 K=1024", 3216, 512, 1024, 0, False, None), ("tc1", 496, 1024, 512, 2, False, None),
]
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = []
for name, M, N, K, act, dot, mdev in SHAPES:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev); dv = torch.randn(N, device=dev)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    tiles = l.fb_gemm_dot_tiles(M, N, K, 1, 0)
    dout = torch.empty(tiles, M, device=dev)
    md = torch.tensor([mdev], dtype=torch.int32, device=dev) if mdev else None
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = act
    if dot:
        g.dotv, g.dot_out, g.dot_stride = dv.data_ptr(), dout.data_ptr(), M
    else:
        g.Cb, g.ldcb = Cb.data_ptr(), N
    g.M, g.N = M, N; g.bf16_mode = 1
    g.m_dev = md.data_ptr() if md is not None else None
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _lib.check(l.fb_gemm(C.byref(g), st), "gemm"); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    Me = mdev or M
    out.append((name, M, N, K, round(ts[len(ts) // 2], 1), round(2 * Me * N * K / ts[len(ts) // 2] / 1e6, 1)))
print("FB_TC_V=" + os.environ.get("FB_TC_V", "2"))
for o in out:
    print(f"{o[0]:16s} M={o[1]:6d} N={o[2]:5d} K={o[3]:5d}  {o[4]:8.1f} us  {o[5]:7.1f} TFLOP/s")
