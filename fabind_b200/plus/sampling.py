"""Batched sampling for FABind+: the reference draws `sample_size` (40) poses per complex by running the
whole dataset `sample_size` times in train() mode (P/test_sampling_fabind.py:126-131, P/inference_sampling_fabind.py:172-185),
i.e. 40 sequential passes over small batches (batch_size 8).  Complexes are independent and the library's dropout masks are keyed
by the ROW of every activation, so S samples of a batch are simply S replicas of its complexes inside ONE larger batch: one
launch sequence, S x more rows per kernel (the latency-bound node-level GEMMs become throughput-bound), a different mask per
replica from one seed.  Only index bookkeeping happens here.
"""
import torch


def _rep_nodes(t, S):
    return torch.cat([t] * S, dim=0)


def _rep_batch(b, S, B):
    return torch.cat([b + k * B for k in range(S)], dim=0)


def _rep_edges(e, S, n_nodes):
    return torch.cat([e + k * n_nodes for k in range(S)], dim=1)


def replicate_batch(data, S, out=None):
    """S tiled copies of the collated batch `data` (replica k of complex i = batch index k * B + i).  Works on any object with the
    reference's HeteroData access pattern (`data['compound'].batch`, `data['complex', 'c2c', 'complex'].edge_index`, ...); `out`
    is an empty object of the same kind to fill (default: fabind_b200.synthetic.HeteroBatch).  Fields read by
    `FABindPlus.inference` only."""
    if out is None:
        from ..synthetic import HeteroBatch
        out = HeteroBatch()
    B = int(data['compound'].batch.max()) + 1
    c, pw, wp = data['compound'], data['protein_whole'], data['complex_whole_protein']
    out['compound'].node_feats = _rep_nodes(c.node_feats, S)
    out['compound'].node_coords = _rep_nodes(c.node_coords, S)
    out['compound'].rdkit_coords = _rep_nodes(c.rdkit_coords, S)
    out['compound'].batch = _rep_batch(c.batch, S, B)
    out['protein_whole'].node_feats = _rep_nodes(pw.node_feats, S)
    out['protein_whole'].batch = _rep_batch(pw.batch, S, B)
    n_wp = wp.batch.shape[0]
    for f in ("node_coords", "node_coords_LAS", "segment", "mask", "is_global"):
        setattr(out['complex_whole_protein'], f, _rep_nodes(getattr(wp, f), S))
    out['complex_whole_protein'].batch = _rep_batch(wp.batch, S, B)
    for rel in ("c2c", "LAS"):
        key = ('complex_whole_protein', rel, 'complex_whole_protein')
        out[key].edge_index = _rep_edges(data[key].edge_index, S, n_wp)
    for name in ("compound_atom_edge_list", "LAS_edge_list"):
        out[name].x = _rep_nodes(data[name].x, S)                      # per-complex local atom indices: unchanged
        out[name].batch = _rep_batch(data[name].batch, S, B)
    out.node_xyz_whole = _rep_nodes(data.node_xyz_whole, S)
    if getattr(data, "pocket_idx", None) is not None:
        out.pocket_idx = _rep_nodes(data.pocket_idx, S)
    if getattr(data, "coords", None) is not None:
        out.coords = _rep_nodes(data.coords, S)
    return out


def sample_batched(model, data, n_samples, seed=0, max_instances=256):
    """`n_samples` poses per complex with ONE `inference` call per chunk of at most `max_instances` (complex, sample) pairs.
    Returns (coords [n_samples, n_atoms, 3], compound_batch [n_atoms], confidence [n_samples, B] or None)."""
    B = int(data['compound'].batch.max()) + 1
    n_atoms = data['compound'].batch.shape[0]
    per = max(1, min(n_samples, max_instances // max(B, 1)))
    flags = {m: m.training for m in model.modules()}   # per-module: sub-modules the caller keeps in eval() stay there
    model.train()
    for name, sub in model.named_modules():
        if name.startswith("confidence") or name.startswith("ranking"):
            sub.eval()
    coords, conf = [], []
    try:
        with torch.no_grad():
            done = 0
            while done < n_samples:
                s = min(per, n_samples - done)
                model.dropout_seed = (int(seed) + 7919 * done) & 0x7FFFFFFF
                res = model.inference(replicate_batch(data, s))
                coords.append(res[0].view(s, n_atoms, 3))
                if len(res) > 2:
                    conf.append(res[2].view(s, B))
                done += s
    finally:
        model.dropout_seed = None
        for m, f in flags.items():
            m.training = f
    return torch.cat(coords, 0), data['compound'].batch, (torch.cat(conf, 0) if conf else None)
