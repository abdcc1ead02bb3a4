"""PDBbind-shaped synthetic complexes (geometry option B).

Produces exactly the tensors `EfficientMCAttModel.forward` takes (reference signature at
FABind/fabind/models/att_model.py:170), laid out as the reference dataloader lays them out
(FABind/fabind/utils/utils.py:328-365): per complex the nodes are
``[glb_c | ligand atoms | glb_p | pocket residues]``; segment 0 = compound side, 1 = protein side;
``mask`` marks the nodes that move between refinement iterations (compound side + glb_p);
coordinates are divided by ``coordinate_scale`` (5 A).

Residues sit on a jittered 5.2 A cubic lattice with a cavity at the origin, the ligand is a random
walk of 1.5 A bonds inside the cavity; bonds = atom pairs closer than 1.7 A (both directions),
LAS (local-atomic-structure) pairs = atom pairs closer than 2.7 A.
"""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class ComplexBatch:
    X: torch.Tensor            # [N, 1, 3] float32, normalised coordinates
    H: torch.Tensor            # [N, embed] float32 node features
    batch_id: torch.Tensor     # [N] int64, sorted
    segment_id: torch.Tensor   # [N] bool (False compound side / True protein side)
    mask: torch.Tensor         # [N] bool, nodes updated between iterations
    is_global: torch.Tensor    # [N] bool
    compound_edge_index: torch.Tensor  # [2, E_bond] int64, global node ids
    LAS_edge_index: torch.Tensor       # [2, E_las] int64
    X_LAS: torch.Tensor        # [N, 1, 3] float32 reference conformer (normalised), zero off-ligand
    n_c: list                  # ligand atoms per complex
    n_p: list                  # pocket residues per complex

    def to(self, device):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.to(device) if torch.is_tensor(v) else v
        return ComplexBatch(**kw)

    def clone(self):
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.clone() if torch.is_tensor(v) else list(v)
        return ComplexBatch(**kw)

    def forward_args(self):
        """Positional/keyword arguments in the reference order (att_model.py:170)."""
        return dict(X=self.X, H=self.H, batch_id=self.batch_id, segment_id=self.segment_id,
                    mask=self.mask, is_global=self.is_global,
                    compound_edge_index=self.compound_edge_index,
                    LAS_edge_index=self.LAS_edge_index,
                    batched_complex_coord_LAS=self.X_LAS, LAS_mask=None)


def _one_complex(rng, n_c, n_p, spacing=5.2, jitter=1.0, cavity=6.0):
    # residues: jittered lattice, cavity around the origin, keep the n_p closest
    half = 2
    while (2 * half + 1) ** 3 < 4 * n_p + 64:
        half += 1
    g = np.arange(-half, half + 1, dtype=np.float64) * spacing
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    pts = pts + rng.uniform(-jitter, jitter, size=pts.shape)
    d = np.linalg.norm(pts, axis=1)
    pts = pts[d > cavity]
    d = d[d > cavity]
    order = np.argsort(d, kind="stable")[:n_p]
    order.sort()
    prot = pts[order]
    # ligand: random walk with 1.5 A steps kept inside the cavity
    lig = np.zeros((n_c, 3))
    for i in range(1, n_c):
        for _ in range(64):
            step = rng.normal(size=3)
            step *= 1.5 / np.linalg.norm(step)
            # branch off a random earlier atom now and then so the graph is not a pure chain
            base = lig[i - 1] if rng.uniform() < 0.8 else lig[rng.integers(0, i)]
            cand = base + step
            if np.linalg.norm(cand) < cavity - 1.0 and (
                    i < 2 or np.min(np.linalg.norm(lig[:i] - cand, axis=1)) > 1.1):
                break
        lig[i] = cand
    dm = np.linalg.norm(lig[:, None] - lig[None], axis=-1)
    np.fill_diagonal(dm, 1e9)
    bonds = np.argwhere(dm < 1.7)
    las = np.argwhere(dm < 2.7)
    lig_ref = lig + rng.normal(scale=0.3, size=lig.shape)
    return prot, lig, bonds, las, lig_ref


def _real_complex(rng, ca, lig_true, bonds, pocket_radius=20.0):
    """One complex from real geometry (SURVEY.md §8d option A): pocket = CA atoms within `pocket_radius` of the ligand
    centroid, coordinates centred on the pocket mean; start pose = the true pose under a random rotation, centred on the
    pocket mean ('redocking' mode, utils/utils.py:321-323); LAS reference = the true pose (:338-346); LAS pairs = bonded
    atoms and atoms two bonds apart."""
    ca, lig_true = np.asarray(ca, np.float64), np.asarray(lig_true, np.float64)
    keep = np.linalg.norm(ca - lig_true.mean(0), axis=1) < pocket_radius
    prot = ca[keep]
    centre = prot.mean(0)
    prot, lig_true = prot - centre, lig_true - centre
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    rot = lig_true @ R.T
    lig = rot - rot.mean(0) + prot.mean(0)
    n = len(lig)
    adj = np.zeros((n, n), dtype=bool)
    adj[bonds[0], bonds[1]] = True
    two = (adj.astype(np.int64) @ adj.astype(np.int64)) > 0
    las = adj | two
    np.fill_diagonal(las, False)
    return prot, lig, np.asarray(bonds).T, np.argwhere(las), lig_true


def make_batch(n_complexes=1, n_c=30, n_p=200, embed=512, seed=0, coordinate_scale=5.0,
               n_c_range=None, n_p_range=None, feature_std=0.1, geometry=None, ids=None):
    """Seeded batch; `n_c_range`/`n_p_range` = (lo, hi) inclusive draw ragged sizes.  `geometry` (a mapping holding
    `<id>_ca`, `<id>_lig`, `<id>_bonds`, e.g. the loaded tests/golden/real_geometry.npz) with `ids` builds the batch from
    real pockets instead of the lattice."""
    rng = np.random.default_rng(seed)
    if geometry is not None:
        n_complexes = len(ids)
    Xs, XL, Hs, bid, seg, msk, glb, bonds_all, las_all = [], [], [], [], [], [], [], [], []
    ncs, nps = [], []
    off = 0
    for b in range(n_complexes):
        nc = int(rng.integers(n_c_range[0], n_c_range[1] + 1)) if n_c_range else n_c
        np_ = int(rng.integers(n_p_range[0], n_p_range[1] + 1)) if n_p_range else n_p
        if geometry is not None:
            prot, lig, bonds, las, lig_ref = _real_complex(rng, geometry[ids[b] + "_ca"], geometry[ids[b] + "_lig"],
                                                           geometry[ids[b] + "_bonds"])
            nc, np_ = len(lig), len(prot)
        else:
            prot, lig, bonds, las, lig_ref = _one_complex(rng, nc, np_)
        n = nc + np_ + 2
        x = np.concatenate([np.zeros((1, 3)), lig, np.zeros((1, 3)), prot], 0)
        xl = np.concatenate([np.zeros((1, 3)), lig_ref, np.zeros((1, 3)), np.zeros_like(prot)], 0)
        Xs.append(x)
        XL.append(xl)
        Hs.append(rng.normal(scale=feature_std, size=(n, embed)))
        bid.append(np.full(n, b))
        s = np.zeros(n, dtype=bool)
        s[nc + 1:] = True
        seg.append(s)
        m = np.zeros(n, dtype=bool)
        m[:nc + 2] = True
        msk.append(m)
        g = np.zeros(n, dtype=bool)
        g[0] = True
        g[nc + 1] = True
        glb.append(g)
        bonds_all.append(bonds.T + 1 + off)
        las_all.append(las.T + 1 + off)
        ncs.append(nc)
        nps.append(np_)
        off += n
    f32 = lambda a: torch.from_numpy(np.concatenate(a, 0)).float()
    X = (f32(Xs) / coordinate_scale).unsqueeze(1).contiguous()
    XLt = (f32(XL) / coordinate_scale).unsqueeze(1).contiguous()
    cat_i = lambda a: torch.from_numpy(np.concatenate(a, 1).astype(np.int64)).contiguous()
    return ComplexBatch(
        X=X, H=f32(Hs).contiguous(),
        batch_id=torch.from_numpy(np.concatenate(bid).astype(np.int64)),
        segment_id=torch.from_numpy(np.concatenate(seg)),
        mask=torch.from_numpy(np.concatenate(msk)),
        is_global=torch.from_numpy(np.concatenate(glb)),
        compound_edge_index=cat_i(bonds_all), LAS_edge_index=cat_i(las_all), X_LAS=XLt,
        n_c=ncs, n_p=nps)


def batch_from_recipe(embed, batch_kwargs, fixture_dir=None):
    """make_batch from a fixture recipe; a `geometry` entry names an .npz under `fixture_dir`."""
    kw = dict(batch_kwargs)
    if isinstance(kw.get("geometry"), str):
        import os
        kw["geometry"] = np.load(os.path.join(fixture_dir, kw["geometry"]))
    return make_batch(embed=embed, **kw)


def randomize_coord_heads(module, std=0.5, seed=1234):
    """The reference initialises every coordinate head with xavier gain 0.001 (egnn.py:52,164), so
    with seeded random weights coordinates barely move.  Parity tests overwrite those heads with
    O(0.1) values so that clamps, the LAS step and the moving inter-edge set are exercised
    (SURVEY.md section 8c)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("coord_mlp.2.weight") or name.endswith("coord_mlp.linear2.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * std / (p.shape[1] ** 0.5) * 4.0)
    return module


# ------------------------------------------------------------------------------------------------------
# Whole-protein docking batches for the L2 wrapper (models/model.py): a duck-typed stand-in for the
# collated torch_geometric HeteroData the reference dataloader produces (utils/utils.py:202-442).
# ------------------------------------------------------------------------------------------------------
class _Store:
    """attribute bag (one node/edge store of a HeteroData)"""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def to(self, device):
        return _Store(**{k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()})


class HeteroBatch:
    """`data['compound'].batch`, `data['complex', 'c2c', 'complex'].edge_index`, `data.pocket_idx` ... as the
    reference's model.forward / model.inference access them."""

    def __init__(self):
        object.__setattr__(self, "_stores", {})

    def __getitem__(self, key):
        if key not in self._stores:
            self._stores[key] = _Store()
        return self._stores[key]

    def to(self, device):
        out = HeteroBatch()
        for k, v in self._stores.items():
            out._stores[k] = v.to(device)
        for k, v in self.__dict__.items():
            if k != "_stores":
                setattr(out, k, v.to(device) if torch.is_tensor(v) else v)
        return out

    def clone(self):
        out = HeteroBatch()
        for k, v in self._stores.items():
            out._stores[k] = _Store(**{a: (t.clone() if torch.is_tensor(t) else t) for a, t in v.__dict__.items()})
        for k, v in self.__dict__.items():
            if k != "_stores":
                setattr(out, k, v.clone() if torch.is_tensor(v) else v)
        return out


def make_docking_batch(n_complexes=2, seed=0, n_c_range=(10, 40), L_range=(150, 400), protein_feat=1280,
                       pocket_radius=20.0):
    """Whole proteins (jittered 5.2 A lattice inside a sphere, centred like utils.py:209-211), a ligand random
    walk near a surface-ish site, torchdrug-like 56-d atom features and ESM-like 1280-d residue features."""
    rng = np.random.default_rng(seed)
    d = HeteroBatch()
    comp_feats, comp_coords, comp_rdkit, comp_batch = [], [], [], []
    prot_feats, prot_xyz, prot_batch, pocket_idx = [], [], [], []
    wp_coords, wp_las, wp_seg, wp_mask, wp_glb, wp_batch, wp_c2c, wp_LAS = [], [], [], [], [], [], [], []
    cx_coords, cx_las, cx_seg, cx_mask, cx_glb, cx_batch, cx_c2c, cx_LAS = [], [], [], [], [], [], [], []
    ael_x, ael_b, lel_x, lel_b = [], [], [], []
    pocket_xyz, pocket_batch, keep_all, dis_map, centers = [], [], [], [], []
    lig_true_all, pocket_centers = [], []
    off_wp = off_cx = 0
    for b in range(n_complexes):
        nc = int(rng.integers(n_c_range[0], n_c_range[1] + 1))
        L = int(rng.integers(L_range[0], L_range[1] + 1))
        half = 2
        while (2 * half + 1) ** 3 < 2 * L + 64:
            half += 1
        gl = np.arange(-half, half + 1, dtype=np.float64) * 5.2
        pts = np.stack(np.meshgrid(gl, gl, gl, indexing="ij"), -1).reshape(-1, 3) + rng.uniform(-1, 1, size=((2 * half + 1) ** 3, 3))
        site = rng.normal(size=3)
        site = site / np.linalg.norm(site) * 9.0                 # binding site 9 A off the protein centre
        keep_pts = np.linalg.norm(pts - site, axis=1) > 6.0        # cavity
        pts = pts[keep_pts]
        order = np.argsort(np.linalg.norm(pts, axis=1), kind="stable")[:L]
        order.sort()
        prot = pts[order]
        prot = prot - prot.mean(0)                                  # utils.py:209-211
        site = site - pts[order].mean(0)
        lig = np.zeros((nc, 3))
        for i in range(1, nc):
            for _ in range(64):
                step = rng.normal(size=3)
                step *= 1.5 / np.linalg.norm(step)
                base = lig[i - 1] if rng.uniform() < 0.8 else lig[rng.integers(0, i)]
                cand = base + step
                if np.linalg.norm(cand) < 5.0 and (i < 2 or np.min(np.linalg.norm(lig[:i] - cand, axis=1)) > 1.1):
                    break
            lig[i] = cand
        lig_true = lig + site                                       # ground-truth pose
        dm = np.linalg.norm(lig[:, None] - lig[None], axis=-1)
        np.fill_diagonal(dm, 1e9)
        bonds = np.argwhere(dm < 1.7)
        las = np.argwhere(dm < 2.7)
        rdkit = lig + rng.normal(scale=0.3, size=lig.shape)        # "rdkit conformer"
        com = lig_true.mean(0)
        keep = np.linalg.norm(prot - com, axis=1) < pocket_radius
        if keep.sum() < 5:
            keep[:100] = True
        pocket = prot[keep]
        coords_init = rdkit - rdkit.mean(0) + pocket.mean(0)        # pocket_center_rdkit, utils.py:319
        z = np.zeros((1, 3))
        n_wp, n_cx = nc + L + 2, nc + int(keep.sum()) + 2
        wp_coords.append(np.concatenate([z, coords_init - coords_init.mean(0), z, prot]))
        wp_las.append(np.concatenate([z, rdkit, z, np.zeros_like(prot)]))
        cx_coords.append(np.concatenate([z, coords_init, z, pocket]))
        cx_las.append(np.concatenate([z, rdkit, z, np.zeros_like(pocket)]))
        for (n, seg_l, msk_l, glb_l, bat_l) in ((n_wp, wp_seg, wp_mask, wp_glb, wp_batch), (n_cx, cx_seg, cx_mask, cx_glb, cx_batch)):
            s = np.zeros(n, dtype=np.float32); s[nc + 1:] = 1
            m = np.zeros(n, dtype=bool); m[:nc + 2] = True
            g = np.zeros(n, dtype=bool); g[0] = True; g[nc + 1] = True
            seg_l.append(s); msk_l.append(m); glb_l.append(g); bat_l.append(np.full(n, b))
        wp_c2c.append(bonds.T + 1 + off_wp); wp_LAS.append(las.T + 1 + off_wp)
        cx_c2c.append(bonds.T + 1 + off_cx); cx_LAS.append(las.T + 1 + off_cx)
        ael_x.append(bonds + 1); ael_b.append(np.full(len(bonds), b))
        lel_x.append(las + 1); lel_b.append(np.full(len(las), b))
        comp_feats.append(rng.normal(scale=0.5, size=(nc, 56))); comp_coords.append(coords_init); comp_rdkit.append(rdkit)
        comp_batch.append(np.full(nc, b))
        prot_feats.append(rng.normal(scale=0.3, size=(L, protein_feat))); prot_xyz.append(prot); prot_batch.append(np.full(L, b))
        pocket_idx.append(keep.astype(np.int32)); keep_all.append(keep)
        pocket_xyz.append(pocket); pocket_batch.append(np.full(int(keep.sum()), b))
        dmap = np.linalg.norm(pocket[:, None] - lig_true[None], axis=-1).reshape(-1)
        dis_map.append(np.minimum(dmap, 10.0)); centers.append(com); lig_true_all.append(lig_true); pocket_centers.append(pocket.mean(0))
        off_wp += n_wp; off_cx += n_cx
    f = lambda a, ax=0: torch.from_numpy(np.concatenate(a, ax)).float()
    li = lambda a, ax=0: torch.from_numpy(np.concatenate(a, ax).astype(np.int64))
    d['compound'].node_feats = f(comp_feats); d['compound'].node_coords = f(comp_coords)
    d['compound'].rdkit_coords = f(comp_rdkit); d['compound'].batch = li(comp_batch)
    d['protein_whole'].node_feats = f(prot_feats); d['protein_whole'].batch = li(prot_batch)
    d['pocket'].batch = li(pocket_batch); d['pocket'].keepNode = torch.from_numpy(np.concatenate(keep_all))
    for name, (co, la, sg, mk, gb, bt, c2c, LAS) in {
            'complex_whole_protein': (wp_coords, wp_las, wp_seg, wp_mask, wp_glb, wp_batch, wp_c2c, wp_LAS),
            'complex': (cx_coords, cx_las, cx_seg, cx_mask, cx_glb, cx_batch, cx_c2c, cx_LAS)}.items():
        d[name].node_coords = f(co); d[name].node_coords_LAS = f(la)
        d[name].segment = torch.from_numpy(np.concatenate(sg)); d[name].mask = torch.from_numpy(np.concatenate(mk))
        d[name].is_global = torch.from_numpy(np.concatenate(gb)); d[name].batch = li(bt)
        d[name, 'c2c', name].edge_index = li(c2c, 1); d[name, 'LAS', name].edge_index = li(LAS, 1)
    d['compound_atom_edge_list'].x = li(ael_x); d['compound_atom_edge_list'].batch = li(ael_b)
    d['LAS_edge_list'].x = li(lel_x); d['LAS_edge_list'].batch = li(lel_b)
    d.node_xyz = f(pocket_xyz); d.node_xyz_whole = f(prot_xyz)
    d.coords_center = torch.from_numpy(np.stack(centers)).float()
    d.pocket_idx = torch.from_numpy(np.concatenate(pocket_idx))
    d.dis_map = f(dis_map)
    d.pocket_residue_center = torch.from_numpy(np.stack(pocket_centers)).float()   # FABind+ dataloader field (P/models/model.py:179)
    d.coords = f(lig_true_all)            # ground-truth ligand pose (FABind+ model.forward shifts it in place, P/models/model.py:257)
    return d
