"""Side benchmarks for the other BASELINE.json configs the round-1 build covers (the graded line is bench.py):
  config 1: single complex (n_c=30, n_p=200), 1 layer x 1 iteration, fp32 parity mode
  config 3: batch=64 with pocket prediction + docking stack (L2 wrapper), ligands 10-80 atoms, bf16
  config 4: FABind+ sampling mode, per-GPU share: batch=32 complexes, one dropout sample per pass (40 passes per complex in the
            reference's protocol), whole FABindPlus.inference (pocket stage + DBSCAN + docking stack + confidence head), bf16
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
WHICH = sys.argv[1:] or ["1", "3", "4"]
# ---- config 1
m1 = EfficientMCAttModel(published_args(), 512, 512, 1, n_layers=1, n_iter=1, normalize_coord=lambda x: x / 5.0,
                         unnormalize_coord=lambda x: x * 5.0)
randomize_coord_heads(m1)
m1 = m1.to(dev).eval()
b1 = make_batch(n_complexes=1, n_c=30, n_p=200, seed=0).to(dev)
X0 = b1.X.clone()
def f1():
    b1.X.copy_(X0); m1(**b1.forward_args())
for prec in (("fp32", "bf16") if "1" in WHICH else ()):
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
ms = timed(f3, steps=5, warmup=2) if "3" in WHICH else float("nan")
nres = int(d3['protein_whole'].batch.shape[0])
if "3" in WHICH:
  print(json.dumps(dict(config="3: batch=64, pocket prediction + docking stack + distance head (L2 wrapper)", dtype="bf16",
                      ms_per_batch=round(ms, 2), complexes_per_s=round(64e3 / ms, 1), residues_total=nres)))

# ---- config 4 (FABind+ sampling mode; random_n_iter off so that every pass does the full 8 iterations)
from fabind_b200.config import published_args_plus
from fabind_b200.plus import FABindPlus, EfficientMCAttModel as PlusStack
import random
if "4" in WHICH:
    a4 = published_args_plus(confidence_training=True, stack_mlp=True, use_clustering=True, random_n_iter=False)
    m4 = FABindPlus(a4, 512, 128)
    randomize_coord_heads(m4)
    m4 = m4.to(dev).train()
    for name, sub in m4.named_modules():
        if name.startswith("confidence") or name.startswith("ranking"):
            sub.eval()
    m4.precision = "bf16"
    d4 = make_docking_batch(32, seed=4, n_c_range=(10, 80), L_range=(150, 800)).to(dev)
    random.seed(0)
    k = [0]
    def f4():
        k[0] += 1
        m4.dropout_seed = k[0]
        with torch.no_grad():
            m4.inference(d4)
    ms = timed(f4, steps=5, warmup=2)
    print(json.dumps(dict(config="4: FABind+ sampling mode, batch=32 x 1 dropout sample per pass (FABindPlus.inference: pocket stage, "
                                 "DBSCAN clustering, 5-layer x 8-iteration docking stack with in-kernel dropout, confidence head)",
                          dtype="bf16", ms_per_pass=round(ms, 2), instances_per_s=round(32e3 / ms, 1),
                          complexes_per_s_at_40_samples=round(32e3 / ms / 40, 2))))
    # the reference's own sampling protocol: batch_size 8, 40 samples per complex (README.md:142-156): 40 sequential passes
    # (sample()) against ONE batched launch sequence per chunk of replicas (plus/sampling.py::sample_batched)
    from fabind_b200.plus.sampling import sample_batched
    d8 = make_docking_batch(8, seed=5, n_c_range=(10, 80), L_range=(150, 800)).to(dev)
    def f_seq():
        m4.sample(lambda: d8, 8, seed=1)
    def f_bat():
        sample_batched(m4, d8, 40, seed=1, max_instances=160)
    ms_seq = timed(f_seq, steps=3, warmup=1) / 8 * 40      # 8 of the 40 passes timed
    ms_bat = timed(f_bat, steps=3, warmup=1)
    print(json.dumps(dict(config="4b: FABind+ sampling, batch_size 8 x 40 samples per complex (reference protocol)", dtype="bf16",
                          ms_40_sequential_passes=round(ms_seq, 1), ms_batched=round(ms_bat, 1),
                          complexes_per_s_sequential=round(8e3 / ms_seq, 2), complexes_per_s_batched=round(8e3 / ms_bat, 2))))
    # the docking stack alone at the config-2 shape, eval vs sampling mode (cost of the in-kernel masks)
    torch.manual_seed(0)
    ms_ = {}
    st4 = PlusStack(published_args_plus(random_n_iter=False), 512, 512, 1, n_layers=5, n_iter=8, normalize_coord=lambda x: x / 5.0,
                    unnormalize_coord=lambda x: x * 5.0)
    randomize_coord_heads(st4)
    st4 = st4.to(dev)
    st4.precision, st4.return_pair = "bf16", False
    b4 = make_batch(n_complexes=16, n_c=30, n_p=200, seed=0).to(dev)
    X4 = b4.X.clone()
    for mode in ("eval", "sampling"):
        st4.train(mode == "sampling")
        def fs():
            b4.X.copy_(X4)
            with torch.no_grad():
                st4(**b4.forward_args())
        ms_[mode] = round(timed(fs, steps=5, warmup=2), 2)
    print(json.dumps(dict(config="FABind+ docking stack, batch=16 (n_c=30, n_p=200), 5 layers x 8 iterations", dtype="bf16",
                          ms_eval=ms_["eval"], ms_sampling=ms_["sampling"], complexes_per_s_eval=round(16e3 / ms_["eval"], 1),
                          instances_per_s_sampling=round(16e3 / ms_["sampling"], 1))))

# ---- ligand post-optimisation (post_optim_utils.py): 64 ligands x 1000 Adam steps in one launch vs the reference's per-ligand CPU loop
if "post" in WHICH or len(sys.argv) == 1:
    import numpy as np
    from fabind_b200.post_optim import post_optimize_batch
    from fabind_b200.synthetic import _one_complex
    refs, preds, batch, las_l, las_b = [], [], [], [], []
    rng = np.random.default_rng(0)
    for b in range(64):
        n = int(rng.integers(10, 81))
        _, lig, _, las, lig_ref = _one_complex(rng, n, 30)
        refs.append(torch.tensor(lig_ref, dtype=torch.float32)); preds.append(torch.tensor(lig + rng.normal(scale=0.8, size=lig.shape), dtype=torch.float32))
        batch.append(torch.full((n,), b)); las_l.append(torch.tensor(las.T.copy())); las_b.append(torch.full((las.shape[0],), b))
    R, P_, Bt, L_, Lb = torch.cat(refs).to(dev), torch.cat(preds).to(dev), torch.cat(batch).to(dev), torch.cat(las_l, 1).to(dev), torch.cat(las_b).to(dev)
    ms = timed(lambda: post_optimize_batch(R, P_, Bt, L_, Lb, total_epoch=1000), steps=3, warmup=1)
    print(json.dumps(dict(config="post-optimisation: 64 ligands (10-80 atoms) x 1000 Adam steps, one launch", ms=round(ms, 2),
                          ligands_per_s=round(64e3 / ms, 1))))
