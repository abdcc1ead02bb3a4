import sys, os, time, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fabind_b200.config import published_args_plus
from fabind_b200.plus import FABindPlus
from fabind_b200.synthetic import make_docking_batch, randomize_coord_heads
dev = "cuda"
for clus in (True, False):
    a4 = published_args_plus(confidence_training=True, stack_mlp=True, use_clustering=clus, random_n_iter=False)
    m4 = FABindPlus(a4, 512, 128); randomize_coord_heads(m4); m4 = m4.to(dev).train(); m4.precision = "bf16"
    d4 = make_docking_batch(32, seed=4, n_c_range=(10, 80), L_range=(150, 800)).to(dev)
    random.seed(0)
    for i in range(3):
        m4.dropout_seed = i
        with torch.no_grad(): m4.inference(d4)
    torch.cuda.synchronize()
    # phase timing
    t0 = time.perf_counter()
    with torch.no_grad():
        m4._drop = m4._sampling_setup()
        s = m4._pocket_stage(d4); torch.cuda.synchronize(); t1 = time.perf_counter()
        m4._cluster_centers(s); torch.cuda.synchronize(); t2 = time.perf_counter()
        Xo, Ho, _ = m4._dock(s, d4, want_pair=False); torch.cuda.synchronize(); t3 = time.perf_counter()
        c = m4._confidence(s, Ho); torch.cuda.synchronize(); t4 = time.perf_counter()
    print(json.dumps(dict(clustering=clus, pocket_ms=round((t1-t0)*1e3,1), cluster_ms=round((t2-t1)*1e3,1), dock_ms=round((t3-t2)*1e3,1), conf_ms=round((t4-t3)*1e3,2),
                          nodes_whole=int(d4['complex_whole_protein'].batch.shape[0]), nP=int(s['nP'].sum()), nA=int(s['nA'].sum()))))
