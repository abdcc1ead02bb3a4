"""fabind_b200.plus.dbscan against sklearn.cluster.DBSCAN (what the reference calls, P/models/model.py:57-61,158)."""
import numpy as np
import pytest

from fabind_b200.plus.dbscan import dbscan_labels

sk = pytest.importorskip("sklearn.cluster")


@pytest.mark.parametrize("seed", range(12))
def test_labels_identical_to_sklearn(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 400))
    k = int(rng.integers(1, 6))
    centers = rng.normal(scale=25.0, size=(k, 3))
    pts = (centers[rng.integers(0, k, n)] + rng.normal(scale=rng.uniform(2.0, 9.0), size=(n, 3))).astype(np.float32)
    for eps, ms in ((9.0, 2), (6.0, 4), (3.0, 1), (12.0, 7)):
        ref = sk.DBSCAN(eps=eps, min_samples=ms).fit(pts).labels_
        assert np.array_equal(dbscan_labels(pts, eps, ms), ref), (seed, eps, ms)


def test_degenerate_inputs():
    assert dbscan_labels(np.zeros((0, 3), np.float32), 9.0, 2).shape == (0,)
    assert dbscan_labels(np.zeros((1, 3), np.float32), 9.0, 2).tolist() == [-1]
    assert dbscan_labels(np.zeros((5, 3), np.float32), 9.0, 2).tolist() == [0] * 5      # padded (0,0,0) rows of the top-50 fallback
