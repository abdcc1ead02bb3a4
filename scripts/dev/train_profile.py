"""Development probe: where one training step (bench.py --config 5, one rank) spends its time.  Wraps the phases of
fabind_b200/train.py with CUDA events AND host clocks (a phase whose host time exceeds its device time is launch-bound), then runs
one more step between cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from fabind_b200 import train, backward as bw, weights, layout, _lib

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
S = bench.setup_train(dev, 0)
rec = {}


def wrap(mod, name, tag=None):
    fn = getattr(mod, name)
    tag = tag or name

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        r = fn(*a, **k)
        e1.record(); t1 = time.perf_counter()
        rec.setdefault(tag, []).append((e0, e1, t1 - t0))
        return r
    setattr(mod, name, w)


wrap(train, "_gpu_prev_coords"); wrap(train, "_gpu_edges"); wrap(train, "build_layout"); wrap(train, "pack_state_dict")
wrap(train, "slot_tensors"); wrap(train, "internal_graph"); wrap(bw, "stack_forward_train_v1"); wrap(bw, "stack_backward_v1")
wrap(train, "arena_grads_to_state_dict"); wrap(train, "_forward_half"); wrap(train, "_backward_half")
wrap(weights.FastPackerV1, "pack", "packer.pack"); wrap(weights.FastPackerV1, "unpack", "packer.unpack")
for _n in ("gcl_forward_train", "att_forward_train", "gcl_backward", "att_backward", "las_bwd", "pair_bias_gate_bwd", "pair_outer_bwd"):
    if hasattr(bw, _n):
        wrap(bw, _n)
for _ in range(3):
    S["step"]()
torch.cuda.synchronize()
rec.clear()
lib = _lib.lib()
l0 = lib.fb_launch_count()
N = 3
t0 = time.perf_counter()
for _ in range(N):
    S["step"]()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / N
out = {k: dict(dev_ms=round(sum(a.elapsed_time(b) for a, b, _ in v) / N, 3), host_ms=round(1e3 * sum(h for _, _, h in v) / N, 3)) for k, v in rec.items()}
print(json.dumps(dict(wall_ms_per_step=round(1e3 * wall, 2), launches_per_step=(lib.fb_launch_count() - l0) / N, phases=out), indent=1))
torch.cuda.profiler.start()
S["step"]()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
