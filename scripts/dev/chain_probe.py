"""Development probe: latency of a CHAIN of dependent node-level GEMMs (each reads the previous output), with the
per-kernel phase timestamps (globaltimer) of sampled CTAs, to see where the time between kernels goes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

def mk(M, N, K, A, Cb, W, b, both=None):
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = 0
    g.Cb, g.ldcb = Cb.data_ptr(), N
    if both is not None:
        g.C, g.ldc = both.data_ptr(), N; g.res, g.ldres = both.data_ptr(), N
    g.M, g.N = M, N; g.bf16_mode = 1
    return g

for (M, N, K, both) in [(3712, 512, 512, False), (3712, 512, 512, True), (496, 512, 512, False), (3712, 1024, 1024, False), (232, 512, 512, False)]:
    bufs = [torch.randn(M, K, device=dev).to(torch.bfloat16) for _ in range(2)]
    W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.zeros(N, device=dev)
    Cf = torch.zeros(M, N, device=dev) if both else None
    gs = [mk(M, N, K, bufs[i & 1], bufs[(i + 1) & 1], W, b, Cf) for i in range(2)]
    n = 64
    for _ in range(2):
        for i in range(n):
            l.fb_gemm(C.byref(gs[i & 1]), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        l.fb_gemm(C.byref(gs[i & 1]), st)
    e1.record(); torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / n * 1e3
    # timeline of 6 consecutive kernels
    dbgs = [torch.zeros(4096, dtype=torch.int64, device=dev) for _ in range(6)]
    for i in range(8):
        l.fb_gemm(C.byref(gs[i & 1]), st)
    for i in range(6):
        l.fb_gemm_set_debug(C.c_void_p(dbgs[i].data_ptr())); l.fb_gemm(C.byref(gs[i & 1]), st)
    torch.cuda.synchronize(); l.fb_gemm_set_debug(None)
    t0 = None
    rows = []
    for i in range(6):
        x = dbgs[i][2048:2048 + 16 * 16].view(-1, 16).cpu().double(); x = x[x[:, 0] > 0]
        d = dbgs[i][:2048].view(-1, 8).cpu().double(); d = d[d[:, 0] > 0]
        if i == 3 and x.shape[0]:
            base = d[:x.shape[0], 4:5]
            print("    epilogue detail (ns after accum ready, mean over sampled CTAs):", [round(v) for v in ((x[:, :13] - base).mean(0)).tolist()])
        if t0 is None:
            t0 = d[:, 0].min().item()
        rows.append(dict(k=i, start=[round(d[:, 0].min().item() - t0), round(d[:, 0].max().item() - t0)],
                         setup=round(d[:, 1].max().item() - t0), tma0=[round(d[:, 2].min().item() - t0), round(d[:, 2].max().item() - t0)],
                         acc0=round(d[:, 4].max().item() - t0), epi0=round(d[:, 5].max().item() - t0), exit=round(d[:, 6].max().item() - t0)))
    print(json.dumps(dict(M=M, N=N, K=K, both=both, chain_us_per_gemm=round(per, 2))))
    for r in rows:
        print("   ", json.dumps(r))
