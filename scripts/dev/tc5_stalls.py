"""Development probe (needs a -DFB_DIAG build: python -m fabind_b200.build --force --diag): who waits for whom inside the multi-problem
node-level GEMM (gemm_tc5.cu) on the launch groups of one layer at the benched size (B = 16: Nc = 496 compound-side rows, Np = 3216
protein-side rows, H = 512).  Clocks per CTA, mean over the CTAs of the launch."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
from fabind_b200 import _lib
l = _lib.lib()
dev = "cuda"
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
H, HD, Nc, Np = 512, 128, 496, 3216
N = Nc + Np
keep = []


def prob(M, Nout, K1, K2=0, act=0, res=False, out="bf16"):
    g = _lib.GemmParams()
    A = torch.randn(M, K1, device=dev).to(torch.bfloat16)
    W = (torch.randn(Nout, K1 + K2, device=dev) / (K1 + K2) ** 0.5).to(torch.bfloat16)
    b = torch.randn(Nout, device=dev)
    g.A, g.lda, g.K1, g.W, g.bias, g.act = A.data_ptr(), K1, K1, W.data_ptr(), b.data_ptr(), act
    keep.extend([A, W, b])
    if K2:
        A2 = torch.randn(M, K2, device=dev).to(torch.bfloat16)
        g.A2, g.lda2, g.K2 = A2.data_ptr(), K2, K2
        keep.append(A2)
    if out in ("f32", "both") or res:
        Cf = torch.randn(M, Nout, device=dev)
        g.C, g.ldc = Cf.data_ptr(), Nout
        keep.append(Cf)
        if res:
            R = torch.randn(M, Nout, device=dev)
            g.res, g.ldres = R.data_ptr(), Nout
            keep.append(R)
    if out in ("bf16", "both"):
        Cb = torch.empty(M, Nout, dtype=torch.bfloat16, device=dev)
        g.Cb, g.ldcb = Cb.data_ptr(), Nout
        keep.append(Cb)
    g.M, g.N, g.bf16_mode = M, Nout, 1
    return g


LDQK = 2 * H + 128
groups = {
    "Pn (first edge-MLP Linear per node), single": [prob(N, 2 * H, H, out="bf16")],
    "n1 (node_mlp.0 on [h | agg], SiLU), single": [prob(N, H, H, H, act=1, out="bf16")],
    "n2 + residual || first projections of the block (3 problems)": [prob(Nc, 4 * HD, H, H, out="f32"), prob(Np, 2 * HD, H, H, out="f32"),
                                                                     prob(N, H, H, res=True, out="both")],
    "L3: TH_p | CAp2 | linear_o(p) + residual": [prob(Np, 2 * H, H, HD, act=2), prob(Np, 2 * HD, H, HD, out="f32"), prob(Np, H, HD, res=True, out="both")],
    "L5: p transition.2 + residual | TH_c | linear_o(c) + residual": [prob(Np, H, 2 * H, res=True, out="both"), prob(Nc, 2 * H, H, HD, act=2),
                                                                      prob(Nc, H, HD, res=True, out="both")],
    "L6: q|k|v(c) | c transition.2 + residual | q|k|v(p)": [prob(Nc, LDQK + 2 * H, H, 2 * H, out="f32"), prob(Nc, H, 2 * H, res=True, out="both"),
                                                            prob(Np, LDQK + 2 * H, H, out="f32")],
}
names = ["mma_wait_operands", "mma_wait_drained_accumulator", "producer0_wait_free_slot", "epilogue0_wait_accumulator", "epilogue0_busy",
         "kernel_total", "wait_previous_grid", "tiles_of_cta"]
for tag, ps in groups.items():
    # the launch orders problems by descending K itself? no: the caller does (forward.cu); mirror it
    ps = sorted(ps, key=lambda g: -(g.K1 + g.K2))
    arr = (_lib.GemmParams * len(ps))(*ps)
    for _ in range(3):
        _lib.check(l.fb_gemm_multi(arr, len(ps), 1, st), "fb_gemm_multi")
    torch.cuda.synchronize()
    dbg = torch.zeros(8192, dtype=torch.int64, device=dev)
    l.fb_gemm_set_debug(C.c_void_p(dbg.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.check(l.fb_gemm_multi(arr, len(ps), 1, st), "fb_gemm_multi")
    e0.record()
    _lib.check(l.fb_gemm_multi(arr, len(ps), 1, st), "fb_gemm_multi")
    e1.record()
    torch.cuda.synchronize()
    l.fb_gemm_set_debug(None)
    tiles = sum(((g.M + 127) // 128) * (g.N // 128) for g in ps)
    slabs = sum(((g.M + 127) // 128) * (g.N // 128) * ((g.K1 + g.K2) // 64) for g in ps)
    flops = sum(2.0 * g.M * g.N * (g.K1 + g.K2) for g in ps)
    n = min(tiles, 148)
    d = dbg[:n * 8].view(n, 8).cpu().double()
    us = e0.elapsed_time(e1) * 1e3
    print(json.dumps(dict(group=tag, tiles=tiles, k_slabs=slabs, gflop=round(flops / 1e9, 2), us_warm_back_to_back=round(us, 1),
                          tflops=round(flops / us / 1e6, 1), mma_clocks_needed_per_cta=round(slabs / n * 128),
                          mean_clk={k: round(d[:, i].mean().item()) for i, k in enumerate(names)},
                          max_clk_kernel_total=round(d[:, 5].max().item()))))
