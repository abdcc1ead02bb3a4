"""Experiment: replay fb_model_forward from a captured CUDA graph vs. enqueueing its ~880 launches on the stream.
Development aid (not the product path): answers whether a per-shape graph cache would lower the latency floor."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fabind_b200 import EfficientMCAttModel, _lib
from fabind_b200.config import published_args
from fabind_b200.runtime import current_stream_ptr
from fabind_b200.synthetic import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
m = EfficientMCAttModel(published_args(), 512, 512, 1, n_layers=4, n_iter=8,
                        normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).cuda().eval()
m.precision = "bf16"
b = make_batch(n_complexes=B, seed=0, n_c=30, n_p=200).to("cuda")
X0 = b.X.clone()
l = _lib.lib()
orig = l.fb_model_forward
saved = {}


def spy(pref, st):
    saved["p"] = _lib.ModelParams.from_buffer_copy(bytes(pref._obj))
    return orig(pref, st)


for _ in range(2):
    b.X.copy_(X0); m(**b.forward_args())
l.fb_model_forward = spy
b.X.copy_(X0)
keep = m(**b.forward_args())          # outputs stay alive: the saved params point at them
l.fb_model_forward = orig
torch.cuda.synchronize()
p = saved["p"]
dev = b.X.device
Xv = b.X.view(-1, 3)
assert p.X_in == Xv.data_ptr(), "inputs were re-staged; the probe needs the caller's buffers"


def direct():
    Xv.copy_(X0.view(-1, 3))
    _lib.check(orig(C.byref(p), current_stream_ptr(dev)), "fb_model_forward")


def timed(fn, n=8):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    return ts


direct(); torch.cuda.synchronize()
ref_H = keep[1].clone()
t_direct = timed(direct)
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.graph(g, stream=s):
    direct()
torch.cuda.synchronize()
t_graph = timed(g.replay)
same = bool(torch.equal(keep[1], ref_H))
print(json.dumps(dict(B=B, direct_ms=t_direct, graph_ms=t_graph, outputs_equal=same)))
