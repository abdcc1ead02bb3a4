"""Ad-hoc device timing of EfficientMCAttModel.forward (development aid, not the graded bench)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fabind_b200.config import published_args, published_args_plus
from fabind_b200 import EfficientMCAttModel
from fabind_b200.synthetic import make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
L = int(sys.argv[3]) if len(sys.argv) > 3 else 4
IT = int(sys.argv[4]) if len(sys.argv) > 4 else 8
flavour = sys.argv[5] if len(sys.argv) > 5 else "v1"
torch.manual_seed(0)
if flavour == "plus":
    from fabind_b200.plus import EfficientMCAttModel as PlusModel
    m = PlusModel(published_args_plus(), 512, 512, 1, n_layers=L, n_iter=IT,
                  normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).cuda().eval()
    m.return_pair = os.environ.get("QB_PAIR", "0") == "1"
else:
    m = EfficientMCAttModel(published_args(), 512, 512, 1, n_layers=L, n_iter=IT,
                            normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0).cuda().eval()
m.precision = prec
b = make_batch(n_complexes=B, seed=0, n_c=30, n_p=200).to("cuda")
X0 = b.X.clone()
for _ in range(2):
    b.X.copy_(X0); m(**b.forward_args())
torch.cuda.synchronize()
ts = []
for _ in range(5):
    b.X.copy_(X0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record(); m(**b.forward_args()); e1.record(); torch.cuda.synchronize(); t1 = time.time()
    ts.append((e0.elapsed_time(e1), (t1 - t0) * 1e3))
print(json.dumps(dict(flavour=flavour, B=B, prec=prec, L=L, IT=IT, gpu_ms=[round(t[0], 2) for t in ts], wall_ms=[round(t[1], 2) for t in ts],
                      complexes_per_s=round(B / (min(t[0] for t in ts) / 1e3), 2), stats=m.last_stats["inter_edges_per_iter"].tolist(),
                      ctx=m.last_stats["ctx_edges"])))
