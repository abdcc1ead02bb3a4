"""A/B of one environment knob of the library on the bench workload (development aid): runs the forward in two
subprocesses (knob = 0 / 1), prints per-forward device times and the largest difference of the outputs.
    python scripts/dev/ab_probe.py FB_PB_FOLD [B]        (any 0/1 environment knob of the library)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    from fabind_b200 import EfficientMCAttModel
    from fabind_b200.config import published_args
    from fabind_b200.synthetic import make_batch
    B, out = int(sys.argv[2]), sys.argv[3]
    torch.manual_seed(0)
    m = EfficientMCAttModel(published_args(), 512, 512, 1, n_layers=4, n_iter=8,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).cuda().eval()
    m.precision = "bf16"
    b = make_batch(n_complexes=B, seed=0, n_c=30, n_p=200).to("cuda")
    X0 = b.X.clone()
    ts = []
    for i in range(8):
        b.X.copy_(X0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); X, H = m(**b.forward_args()); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    torch.save(dict(X=X.cpu(), H=H.cpu()), out)
    print(json.dumps(ts[2:]))
    sys.exit(0)
knob, B = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "16")
res = {}
for v in ("0", "1"):
    out = f"/tmp/ab_{knob}_{v}.pt"
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", B, out], env=dict(os.environ, **{knob: v}),
                       capture_output=True, text=True, timeout=300)
    res[v] = r.stdout.strip().splitlines()[-1] if r.returncode == 0 else ("FAILED " + r.stderr[-400:])
import torch
try:
    a, b = torch.load(f"/tmp/ab_{knob}_0.pt"), torch.load(f"/tmp/ab_{knob}_1.pt")
    diff = dict(X=float((a["X"] - b["X"]).abs().max()), H=float((a["H"] - b["H"]).abs().max()), Hscale=float(a["H"].abs().max()))
except Exception as e:
    diff = str(e)
print(json.dumps(dict(knob=knob, B=int(B), ms_off=res["0"], ms_on=res["1"], max_abs_diff=diff)))
