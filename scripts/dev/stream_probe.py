"""Experiment: throughput when the batch of 16 is split into S sub-batches issued on S streams by S host threads."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from fabind_b200.synthetic import make_batch
from fabind_b200 import runtime

dev = torch.device("cuda", 0)
model = bench.build_model(dev, "bf16")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def run(S, steps=8, threads=True):
    per = 16 // S
    subs = [make_batch(n_complexes=per, n_c=30, n_p=200, embed=512, seed=100 + i).to(dev) for i in range(S)]
    X0 = [b.X.clone() for b in subs]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    def work(i):
        with torch.cuda.stream(streams[i]):
            subs[i].X.copy_(X0[i])
            model(**subs[i].forward_args())
    def step():
        main = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(main)
        if threads and S > 1:
            ts = [threading.Thread(target=work, args=(i,)) for i in range(S)]
            [t.start() for t in ts]; [t.join() for t in ts]
        else:
            for i in range(S):
                work(i)
        for s in streams:
            main.wait_stream(s)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    tot = 0.0; host = 0.0
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter()
        step()
        host += time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"S={S} threads={threads}: {tot / steps:.3f} ms/step device, host enqueue {host / steps * 1e3:.3f} ms -> {16 / (tot / steps) * 1e3:.0f} complexes/s", flush=True)

# per-thread scratch
_orig = runtime._scratch_buf
def _sb(device, tag, nbytes):
    return _orig(device, (tag, torch.cuda.current_stream(device).cuda_stream), nbytes)
runtime._scratch_buf = _sb
for S, th in ((1, False), (2, False), (2, True), (4, False), (4, True)):
    run(S, threads=th)
