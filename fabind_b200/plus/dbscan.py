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
from scipy.spatial import cKDTree


def dbscan_labels(points, eps, min_samples):
    x = np.ascontiguousarray(points, dtype=np.float64)
    n = x.shape[0]
    labels = np.full(n, -1, dtype=np.int64)
    if n == 0:
        return labels
    # neighbour pairs i < j with distance <= eps (radius_neighbors semantics; the point itself counts towards min_samples)
    pairs = cKDTree(x).query_pairs(float(eps), output_type="ndarray")
    deg = np.bincount(pairs.ravel(), minlength=n) + 1
    core = deg >= min_samples
    if not core.any():
        return labels
    cc = pairs[core[pairs[:, 0]] & core[pairs[:, 1]]]
    g = csr_matrix((np.ones(len(cc), dtype=np.int8), (cc[:, 0], cc[:, 1])), shape=(n, n))
    _, comp = connected_components(g, directed=False)
    idx = np.nonzero(core)[0]
    comp_c = comp[idx]
    # number the components of core points in order of their lowest-index core point (sklearn's discovery order)
    uniq, first = np.unique(comp_c, return_index=True)
    rank = np.empty(comp.max() + 1, dtype=np.int64)
    rank[uniq[np.argsort(first, kind="stable")]] = np.arange(uniq.size)
    labels[idx] = rank[comp_c]
    # border points: reached first by the earliest-expanded (lowest-numbered) cluster among their core neighbours
    big = np.iinfo(np.int64).max
    best = np.full(n, big, dtype=np.int64)
    for a_, b_ in ((0, 1), (1, 0)):
        sel = ~core[pairs[:, a_]] & core[pairs[:, b_]]
        np.minimum.at(best, pairs[sel, a_], labels[pairs[sel, b_]])
    hit = ~core & (best != big)
    labels[hit] = best[hit]
    return labels
