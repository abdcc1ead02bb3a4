"""Development probe: per-phase timeline of the tcgen05 GEMM (globaltimer stamps) + event-timed duration."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import ctypes as C
import torch
from fabind_b200 import _lib
from test_gpu_gemm import run_gemm

l = _lib.lib()
dev = "cuda"
for (M, N, K, act) in [(3712, 512, 512, 0), (44922, 512, 512, 1), (496, 512, 512, 0), (100000, 512, 512, 0), (3712, 1024, 512, 2)]:
    torch.manual_seed(0)
    A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
    ntile = (M + 127) // 128
    dbg = torch.zeros(((ntile + 7) // 8) * 8, dtype=torch.int64, device=dev)
    for _ in range(3):
        run_gemm(A, W, b, act, bf16=True, want_cb=True)
    l.fb_gemm_set_debug(C.c_void_p(dbg.data_ptr()))
    run_gemm(A, W, b, act, bf16=True, want_cb=True)
    l.fb_gemm_set_debug(None)
    d = dbg.view(-1, 8).cpu().double()
    d = d[d[:, 0] > 0]
    rel = (d[:, 1:7] - d[:, 0:1])
    names = ["alloc+sync", "first TMA landed", "all MMA issued", "accum ready (epi)", "epilogue done", "exit sync"]
    span = (d[:, 6].max() - d[:, 0].min()).item()
    # event timing without debug
    Ad, Wd = A.to(torch.bfloat16), W.to(torch.bfloat16)
    ts = []
    g = _lib.GemmParams()
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    g.A, g.lda, g.K1 = Ad.data_ptr(), K, K; g.W = Wd.data_ptr(); g.bias = b.data_ptr(); g.act = act
    g.Cb, g.ldcb = Cb.data_ptr(), N; g.M, g.N = M, N; g.bf16_mode = 1
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for i in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); l.fb_gemm(C.byref(g), st); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(json.dumps(dict(M=M, N=N, K=K, act=act, us_median=round(ts[10], 1), us_min=round(ts[0], 1),
                          tflops=round(2 * M * N * K / ts[10] / 1e6, 1), ctas_sampled=int(d.shape[0]),
                          kernel_span_us=round(span / 1e3, 1),
                          phase_mean_ns={n: round(rel[:, i].mean().item()) for i, n in enumerate(names)},
                          phase_max_ns={n: round(rel[:, i].max().item()) for i, n in enumerate(names)})))
