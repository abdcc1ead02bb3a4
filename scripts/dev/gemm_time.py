"""Development probe: wall time of single fb_gemm launches (bf16 mode) over the shapes of the benched step, CUDA events over REPS
back-to-back launches (operands L2-warm where they fit) and, with --flush, with a 256 MiB memset between launches (timed per launch)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
REPS = 30
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [  # M, N, K, act, dot, out ("bf16" | "f32" | "none" | "both+res")
    (44904, 512, 512, 1, False, "bf16"), (44904, 512, 512, 1, True, "none"), (44904, 512, 512, 0, False, "bf16"),
    (11000, 1024, 576, 2, True, "none"), (99712, 512, 512, 0, False, "bf16"),
    (3712, 1024, 512, 0, False, "bf16"), (3712, 512, 512, 1, False, "bf16"), (3712, 512, 512, 0, False, "both+res"),
    (3712, 2688, 512, 0, False, "bf16"), (3216, 512, 128, 0, False, "both+res"), (496, 512, 512, 0, False, "bf16"),
]
for (M, N, K, act, dot, out) in SHAPES:
    A = torch.randn(M, K, device=dev).to(torch.bfloat16); W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device=dev); dv = torch.randn(N, device=dev)
    Cb = torch.empty(M, N, dtype=torch.bfloat16, device=dev); Cf = torch.randn(M, N, device=dev)
    nt = l.fb_gemm_dot_tiles(M, N, K, 1, 0)
    dout = torch.zeros(nt * M, device=dev)
    g = _lib.GemmParams()
    g.A, g.lda, g.K1 = A.data_ptr(), K, K; g.W = W.data_ptr(); g.bias = b.data_ptr(); g.act = act
    if out in ("bf16", "both+res"):
        g.Cb, g.ldcb = Cb.data_ptr(), N
    if out in ("f32", "both+res"):
        g.C, g.ldc = Cf.data_ptr(), N
    if out == "both+res":
        g.res, g.ldres = Cf.data_ptr(), N
    if dot:
        g.dotv = dv.data_ptr(); g.dot_out = dout.data_ptr(); g.dot_stride = M
    g.M, g.N = M, N; g.bf16_mode = 1
    for _ in range(5):
        _lib.check(l.fb_gemm(C.byref(g), st), "fb_gemm")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        l.fb_gemm(C.byref(g), st)
    e1.record(); torch.cuda.synchronize()
    warm = e0.elapsed_time(e1) / REPS * 1e3
    cold = []
    for _ in range(8):
        flush.zero_()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); l.fb_gemm(C.byref(g), st); z.record(); torch.cuda.synchronize()
        cold.append(a.elapsed_time(z) * 1e3)
    cold.sort()
    fl = 2.0 * M * N * K
    # correctness spot check against torch (bf16 operands, fp32 accumulate)
    ref = A.float() @ W.float().t() + b
    ref = torch.nn.functional.silu(ref) if act == 1 else torch.relu(ref) if act == 2 else ref
    err = None
    if out == "bf16":
        err = ((Cb.float() - ref).abs().max() / ref.abs().max()).item()
    if dot:
        got = dout.view(nt, M).sum(0)
        err = ((got - ref @ dv).abs().max() / (ref @ dv).abs().max()).item()
    print(json.dumps(dict(M=M, N=N, K=K, act=act, dot=dot, out=out, us_warm=round(warm, 2), us_cold_med=round(cold[len(cold) // 2], 2),
                          tflops_warm=round(fl / warm / 1e6, 1), rel_err=err)))
