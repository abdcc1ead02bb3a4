"""Ligand post-optimisation: CPU oracle against goldens from the unmodified reference function (anywhere), GPU kernel against
the oracle and the goldens (-m gpu).  Tolerances: over 10 Adam steps the trajectories agree to ~1e-4 A; beyond that the
objective's sign() terms make ANY two fp32 evaluations diverge (the reference's own result changes with the ligand's absolute
position once torch.cdist switches to the |a|^2+|b|^2-2ab form above 25 atoms; measured 0.01-0.7 A between the reference and a
float32 restatement with identical update rules), and atoms outside every LAS pair are not pinned by the objective at all.  The
fixed-step Adam on an |.| objective never settles below its step (lr = 0.1 A): the iterate jitters by ~0.1 A per step, so the
"last loss" is itself a noisy read-out.  The longer cases are therefore compared on what the optimisation is FOR: final loss
within 0.02 A per constrained pair, RMSD to the conformer within 0.05 A, and the constrained pair distances within 0.15 A on
average (the jitter floor)."""
import os

import pytest
import torch

from oracle.post_optim_oracle import post_optimize
from helpers import GOLDEN_DIR

CASES = torch.load(os.path.join(GOLDEN_DIR, "postopt_cases.pt"), map_location="cpu", weights_only=False)["cases"]


def _check(c, x, loss, rmsd):
    if c["epochs"] <= 10:
        assert float((x - c["x"]).abs().max()) < 5e-4, (c["n"], c["epochs"])
        assert abs(loss - c["loss"]) <= 1e-4 * abs(c["loss"]) and abs(rmsd - c["rmsd"]) < 1e-4
        return
    i, j = (c["las"][0], c["las"][1]) if not c["rigid"] else torch.triu_indices(c["n"], c["n"], 1)
    n_terms = i.numel() * (2 if c["rigid"] else 1)
    assert abs(loss - c["loss"]) <= 0.02 * n_terms, (loss, c["loss"], n_terms)
    assert abs(rmsd - c["rmsd"]) < 0.05, (rmsd, c["rmsd"])
    d_mine, d_ref = (x[i] - x[j]).norm(dim=-1), (c["x"][i] - c["x"][j]).norm(dim=-1)
    assert float((d_mine - d_ref).abs().mean()) < 0.15, float((d_mine - d_ref).abs().mean())


@pytest.mark.parametrize("k", range(len(CASES)))
def test_oracle_matches_reference_golden(k):
    c = CASES[k]
    x, loss, rmsd = post_optimize(c["ref"], c["pred"], c["epochs"], None if c["rigid"] else c["las"])
    _check(c, x, loss, rmsd)


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(len(CASES)))
def test_gpu_matches_reference_golden(k):
    from fabind_b200.post_optim import post_optimize_compound_coords
    c = CASES[k]
    x, loss, rmsd = post_optimize_compound_coords(c["ref"].cuda(), c["pred"].cuda(), c["epochs"], None if c["rigid"] else c["las"].cuda())
    _check(c, x.cpu(), loss, rmsd)


@pytest.mark.gpu
def test_gpu_batch_matches_oracle_short_horizon():
    """a ragged batch of ligands in ONE launch, 30 steps: every ligand against the oracle run on it alone"""
    import numpy as np
    from fabind_b200.post_optim import post_optimize_batch
    from fabind_b200.synthetic import _one_complex
    refs, preds, batch, las_l, las_b = [], [], [], [], []
    for b, n in enumerate((7, 33, 64, 90, 12)):
        rng = np.random.default_rng(b)
        _, lig, _, las, lig_ref = _one_complex(rng, n, 30)
        refs.append(torch.tensor(lig_ref, dtype=torch.float32)); preds.append(torch.tensor(lig + rng.normal(scale=0.6, size=lig.shape), dtype=torch.float32))
        batch.append(torch.full((n,), b)); las_l.append(torch.tensor(las.T.copy())); las_b.append(torch.full((las.shape[0],), b))
    x, loss, rmsd = post_optimize_batch(torch.cat(refs).cuda(), torch.cat(preds).cuda(), torch.cat(batch).cuda(),
                                        torch.cat(las_l, 1).cuda(), torch.cat(las_b).cuda(), total_epoch=30)
    o = 0
    for b, (r, p, l) in enumerate(zip(refs, preds, las_l)):
        xo, lo, ro = post_optimize(r, p, 30, l)
        n = r.shape[0]
        assert float((x[o:o + n].cpu() - xo).abs().max()) < 2e-3, b
        assert abs(float(loss[b]) - lo) < 1e-3 * abs(lo) and abs(float(rmsd[b]) - ro) < 1e-4
        o += n
