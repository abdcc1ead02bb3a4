"""Side benchmarks for the other BASELINE.json configs the round-1 build covers (the graded line is bench.py):
  config 1: single complex (n_c=30, n_p=200), 1 layer x 1 iteration, fp32 parity mode
  config 3: batch=64 with pocket prediction + docking stack (L2 wrapper), ligands 10-80 atoms, bf16
Each prints one JSON line with complexes/s on the device (CUDA events, L2 flushed between steps)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fabind_b200 import EfficientMCAttModel
from fabind_b200.config import published_args
from fabind_b200.model import IaBNet_mean_and_pocket_prediction_cls_coords_dependent as Net
from fabind_b200.synthetic import make_batch, make_docking_batch, randomize_coord_heads

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


torch.manual_seed(0)
# ---- config 1
m1 = EfficientMCAttModel(published_args(), 512, 512, 1, n_layers=1, n_iter=1, normalize_coord=lambda x: x / 5.0,
                         unnormalize_coord=lambda x: x * 5.0)
randomize_coord_heads(m1)
m1 = m1.to(dev).eval()
b1 = make_batch(n_complexes=1, n_c=30, n_p=200, seed=0).to(dev)
X0 = b1.X.clone()
def f1():
    b1.X.copy_(X0); m1(**b1.forward_args())
for prec in ("fp32", "bf16"):
    m1.precision = prec
    ms = timed(f1)
    print(json.dumps(dict(config="1: single complex, 1 layer x 1 iteration", dtype=prec, ms_per_forward=round(ms, 3),
                          complexes_per_s=round(1e3 / ms, 1))))
# ---- config 3
args = published_args()
m3 = Net(args, 512, 128)
randomize_coord_heads(m3)
m3 = m3.to(dev).eval()
m3.precision = "bf16"
d3 = make_docking_batch(64, seed=3, n_c_range=(10, 80), L_range=(150, 800)).to(dev)
def f3():
    m3(d3, stage=2)
ms = timed(f3, steps=5, warmup=2)
nres = int(d3['protein_whole'].batch.shape[0])
print(json.dumps(dict(config="3: batch=64, pocket prediction + docking stack + distance head (L2 wrapper)", dtype="bf16",
                      ms_per_batch=round(ms, 2), complexes_per_s=round(64e3 / ms, 1), residues_total=nres)))
