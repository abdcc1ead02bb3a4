"""Development probe: where one pass of config 3 (L2 wrapper, stage 2, B = 64) spends its time (CUDA-synchronised wall clock per phase)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
dev = torch.device("cuda", 0)
S = bench.setup_l2(dev, 0, 3)
m = S["model"]
rec = {}
def wrap(obj, name):
    fn = getattr(obj, name)
    def w(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); rec.setdefault(name, []).append(time.perf_counter() - t0)
        return r
    setattr(obj, name, w)
for n in ("_pocket_stage", "_centers", "_dock"):
    wrap(m, n)
wrap(m.pocket_pred_model, "forward"); wrap(m.complex_model, "forward")
for _ in range(3):
    S["step"]()
rec.clear()
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 5
for _ in range(N):
    S["step"]()
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / N
print(json.dumps(dict(ms_total=round(tot * 1e3, 2), phases_ms={k: round(1e3 * sum(v) / N, 2) for k, v in rec.items()})))
