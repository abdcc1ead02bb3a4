"""BASELINE config 5 side benchmark: one training step through the v1 docking stack (forward + backward + gradient all-reduce),
16 complexes per GPU (global batch 16 x world), fp32 reverse pass.  Development aid for the training path -- NOT the graded line
(bench.py stays on config 2).  Phases are timed with CUDA events on the launching stream, max over ranks.

    python scripts/bench_train.py                       # 1 GPU
    torchrun --nproc-per-node 8 scripts/bench_train.py  # 8 x 16 = 128 complexes per step

What the step contains (fabind_b200/train.py): iterations 0..n-2 through the inference path (bf16 tcgen05 if --precision bf16),
the training-mode forward of the last iteration, the reverse pass, the packer's chain rule to `state_dict`-shaped gradients and ONE
flat all-reduce.  `iter_i` is fixed to n_iter (the reference draws randint(1, n_iter), att_model.py:210-211); dropout off; the
optimizer step is excluded."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--iters", type=int, default=8)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"], help="precision of the no_grad iterations")
    ap.add_argument("--flavour", default="v1", choices=["v1", "plus"], help="weight layout: FABind (4 layers) or FABind+ (5 layers, LayerNorm MLPs)")
    ap.add_argument("--gemm", default="fp32", choices=["fp32", "bf16"], help="GEMMs of the differentiated iteration (backward.PRECISION)")
    a = ap.parse_args()
    from fabind_b200 import EfficientMCAttModel, train, backward
    backward.PRECISION = a.gemm
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_batch, randomize_coord_heads
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    if a.flavour == "plus":
        from fabind_b200.config import published_args_plus
        from fabind_b200.plus import EfficientMCAttModel as PlusModel
        m = PlusModel(published_args_plus(), a.hidden, a.hidden, 1, n_layers=a.layers, n_iter=a.iters,
                      normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        m.return_pair = False
    else:
        m = EfficientMCAttModel(published_args(), a.hidden, a.hidden, 1, n_layers=a.layers, n_iter=a.iters,
                                normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
        randomize_coord_heads(m, std=0.5)
    m = m.to(dev).eval()
    m.precision = a.precision
    b = make_batch(n_complexes=a.batch, n_c=30, n_p=200, embed=a.hidden, seed=100 + rank).to(dev)
    fa = b.forward_args()
    X0 = b.X.clone()
    g = torch.Generator(device="cpu").manual_seed(1)
    rx, rh = torch.randn(b.X.shape, generator=g).to(dev), torch.randn(b.H.shape, generator=g).to(dev)
    sd = {k: v.detach() for k, v in m.state_dict().items()}

    def step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        fa["X"] = X0.clone()
        ev[0].record()
        if a.flavour == "plus":
            pg = train.training_step(m, fa, lambda X_, H_, P_: (rx, rh, torch.zeros_like(P_)), state_dict=sd)[3]
        else:
            pg = train.training_step(m, fa, lambda X_, H_: (rx, rh), state_dict=sd)[2]
        ev[1].record()
        train.apply_gradients(m, pg)
        ev[2].record()
        torch.cuda.synchronize(dev)
        return ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    ts = [step() for _ in range(a.steps)]
    t = torch.tensor([sum(x[0] for x in ts) / a.steps, sum(x[1] for x in ts) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.sum())
        print(json.dumps(dict(metric=f"training step (fwd + bwd + gradient all-reduce), {a.flavour} stack", n_gpus=world, global_batch=a.batch * world,
                              ms_step=round(ms, 2), ms_forward_backward=round(float(t[0]), 2), ms_grads_allreduce=round(float(t[1]), 2),
                              complexes_per_s=round(a.batch * world / (ms / 1e3), 1), hidden=a.hidden, layers=a.layers, iters=a.iters,
                              no_grad_iterations=a.precision, reverse_pass="fp32 SIMT (first correct version)" if a.gemm == "fp32" else "tcgen05 GEMMs (bf16 operands), SIMT scatter kernels",
                              note="host-side packing / chain rule of the weight arena is inside ms_forward_backward")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
