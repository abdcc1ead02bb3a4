"""Edge cases of the docking stack on the GPU against the CPU oracles (both weight layouts): degenerate ligands (one atom: no
bond, no LAS pair; two atoms), a tiny pocket, a whole-protein-sized complex at the pocket-stage width (1500 residues), and
batch invariance (a complex gives the same result alone and inside a ragged batch - complexes are independent units)."""
import pytest
import torch

from oracle import fabind_oracle as orc
from oracle import fabind_plus_oracle as orcp
from oracle import ref_shims
from oracle.det_weights import det_state_dict
from fabind_b200 import EfficientMCAttModel as V1Model
from fabind_b200.plus import EfficientMCAttModel as PlusModel
from fabind_b200.synthetic import make_batch
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _build(flavour, hidden, L, IT, seed=41):
    cls, args = (PlusModel, ref_shims.published_args_plus()) if flavour == "plus" else (V1Model, ref_shims.published_args())
    m = cls(args, hidden, hidden, 1, n_layers=L, n_iter=IT, normalize_coord=lambda x: x / 5.0, unnormalize_coord=lambda x: x * 5.0)
    sd = det_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval(), sd


def _oracle(flavour, sd, L, IT, b):
    f = orcp.model_forward if flavour == "plus" else orc.model_forward
    with torch.no_grad():
        out = f(sd, orc.make_cfg(n_layers=L, n_iter=IT), b.X, b.H, b.batch_id, b.segment_id, b.mask, b.is_global,
                b.compound_edge_index, b.LAS_edge_index, b.X_LAS)
    return out[0], out[1]


@pytest.mark.parametrize("flavour", ["v1", "plus"])
@pytest.mark.parametrize("bkw", [
    dict(n_complexes=1, seed=1, n_c=1, n_p=10),                                   # single-atom ligand: no bonds, no LAS pairs
    dict(n_complexes=2, seed=2, n_c=2, n_p=6),                                    # two atoms, six residues
    dict(n_complexes=3, seed=3, n_c_range=(1, 4), n_p_range=(3, 12)),             # ragged and tiny
], ids=["one_atom", "two_atoms", "tiny_ragged"])
def test_degenerate_complexes(flavour, bkw):
    hidden, L, IT = 64, 2, 2
    m, sd = _build(flavour, hidden, L, IT)
    b = make_batch(embed=hidden, **bkw)
    Xo, Ho = _oracle(flavour, sd, L, IT, b)
    out = m(**b.to("cuda").forward_args())
    assert rel_err(out[0], Xo) < 1e-4 and rel_err(out[1], Ho) < 1e-4


@pytest.mark.parametrize("flavour", ["v1", "plus"])
def test_whole_protein_size(flavour):
    """pocket-stage shape at its upper end: 1500 residues, 80 ligand atoms, hidden 128, 1 layer x 1 iteration"""
    hidden, L, IT = 128, 1, 1
    m, sd = _build(flavour, hidden, L, IT)
    b = make_batch(embed=hidden, n_complexes=1, seed=4, n_c=80, n_p=1500)
    Xo, Ho = _oracle(flavour, sd, L, IT, b)
    out = m(**b.to("cuda").forward_args())
    assert rel_err(out[0], Xo) < 1e-4 and rel_err(out[1], Ho) < 1e-4


@pytest.mark.parametrize("flavour,precision", [("v1", "fp32"), ("plus", "fp32"), ("v1", "bf16"), ("plus", "bf16")])
def test_batch_invariance(flavour, precision):
    """complex 1 of a ragged batch of 4 == the same complex alone (same kernels, different row offsets / tile positions)"""
    from fabind_b200 import shard
    hidden, L, IT = 128, 2, 2
    m, sd = _build(flavour, hidden, L, IT)
    m.precision = precision
    b = make_batch(embed=hidden, n_complexes=4, seed=6, n_c_range=(6, 30), n_p_range=(30, 120)).to("cuda")
    fa = b.forward_args()
    sub, idx = shard.take_complexes({k: (v.clone() if torch.is_tensor(v) else v) for k, v in fa.items()}, [1])
    alone = m(**sub)
    full = m(**fa)
    tol = 1e-5 if precision == "fp32" else 2e-2      # bf16: tile boundaries move, accumulation order inside a GEMM tile does not
    assert rel_err(alone[0], full[0][idx]) < tol and rel_err(alone[1], full[1][idx]) < tol
