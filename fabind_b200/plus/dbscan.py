"""DBSCAN cluster labels for the FABind+ pocket-centre clustering (`--use-clustering`, P/models/model.py:147-167).

The reference calls `sklearn.cluster.DBSCAN(eps, min_samples).fit(points)` once per complex on the host and then picks a
cluster with python's `random`; sklearn's per-call overhead (input validation, neighbour-tree construction, ~2 ms for a few
hundred points) is 40 % of a sampling pass once the model runs on the GPU.  This is the same algorithm restated for one small
point set (dense distance matrix, connected components of the core graph, border points to the first-discovered cluster),
returning labels IDENTICAL to sklearn's (numbering = order of discovery by lowest point index; -1 = noise); checked against
sklearn on random inputs in tests/test_dbscan.py.  Host logic on both sides of the comparison.
"""
import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components


def dbscan_labels(points, eps, min_samples):
    x = np.asarray(points, dtype=np.float64)
    n = x.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    if n == 0:
        return labels
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    adj = d2 <= float(eps) ** 2                     # radius_neighbors: distance <= eps, the point itself included
    core = adj.sum(1) >= min_samples
    idx = np.nonzero(core)[0]
    if idx.size == 0:
        return labels
    _, comp = connected_components(csr_matrix(adj[np.ix_(idx, idx)]), directed=False)
    # number the components in order of their lowest-index core point (sklearn's discovery order)
    first = np.full(comp.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, comp, np.arange(idx.size))
    rank = np.empty_like(first)
    rank[np.argsort(first, kind="stable")] = np.arange(first.size)
    labels[idx] = rank[comp]
    # border points: reached first by the earliest-expanded (lowest-numbered) cluster among their core neighbours
    border = np.nonzero(~core & (adj & core[None, :]).any(1))[0]
    for b in border:
        labels[b] = labels[np.nonzero(adj[b] & core)[0]].min()
    return labels
