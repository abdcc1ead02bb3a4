"""Import the UNMODIFIED reference model files from /root/reference (dev container only).

TEST INFRASTRUCTURE -- only tests/, scripts/make_golden.py and the validation of oracle/ use this.
Nothing here is importable on the GPU box (the reference tree does not travel with the repo).

The reference needs two third-party packages that are not installed here:
  * torch_scatter 2.1.0  (FABind/README.md:40)  -- call sites egnn.py:13,221,444,777; att_model.py:7,43
  * torch_geometric 2.4.0 (FABind/README.md:44) -- to_dense_batch at egnn.py:264-265, att_model.py:203-204
Their published semantics are restated below in plain torch (index_add / segment softmax / stable
scatter into a padded [B, max_n, ...] block) and injected through sys.modules so that the reference
files import without being edited.
"""
import os
import sys
import types
import importlib

import torch

REF_ROOT = "/root/reference"
REF_V1 = os.path.join(REF_ROOT, "FABind", "fabind")
REF_PLUS = os.path.join(REF_ROOT, "FABind_plus", "fabind")


def reference_available():
    return os.path.isdir(os.path.join(REF_V1, "models"))


# ---- torch_scatter restatement (documented semantics of pytorch_scatter 2.1.0) -----------------
def _scatter_sum(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0, "the reference only scatters along dim 0"
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
    res.index_add_(0, index, src)
    return res


def _scatter_softmax(src, index, dim=0, dim_size=None):
    # max-shifted exp divided by the segment sum (pytorch_scatter composite/softmax.py)
    n = int(index.max()) + 1 if dim_size is None else dim_size
    mx = src.new_full((n,), float("-inf")).scatter_reduce(0, index, src, reduce="amax", include_self=True)
    e = (src - mx[index]).exp()
    s = src.new_zeros((n,)).index_add_(0, index, e)
    return e / s[index]


def _scatter_max(src, index, dim=0, dim_size=None):  # imported (unused) by FABind+ model.py:11
    n = int(index.max()) + 1 if dim_size is None else dim_size
    mx = src.new_full((n,) + tuple(src.shape[1:]), float("-inf"))
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return mx.scatter_reduce(0, idx, src, reduce="amax", include_self=True), None


def _scatter_mean(src, index, dim=0, dim_size=None):
    s = _scatter_sum(src, index, dim, dim_size=dim_size)
    c = _scatter_sum(torch.ones_like(src), index, dim, dim_size=dim_size)
    return s / c.clamp(min=1)


# ---- torch_geometric.utils restatement ----------------------------------------------------------
def _to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    """Stable scatter of rows of `x` (sorted by `batch`) into [B, max_n, ...] + bool mask."""
    if batch is None:
        return x.unsqueeze(0), x.new_ones((1, x.shape[0]), dtype=torch.bool)
    B = int(batch.max()) + 1 if batch_size is None else batch_size
    counts = torch.zeros(B, dtype=torch.long, device=x.device).index_add_(0, batch, torch.ones_like(batch))
    starts = torch.cumsum(counts, 0) - counts
    max_n = int(counts.max()) if max_num_nodes is None else max_num_nodes
    pos = torch.arange(x.shape[0], device=x.device) - starts[batch]
    out = x.new_full((B, max_n) + tuple(x.shape[1:]), fill_value)
    out[batch, pos] = x
    mask = torch.zeros((B, max_n), dtype=torch.bool, device=x.device)
    mask[batch, pos] = True
    return out, mask


def _to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None):  # post-optim only
    n = int(edge_index.max()) + 1 if max_num_nodes is None else max_num_nodes
    adj = torch.zeros((1, n, n))
    adj[0, edge_index[0], edge_index[1]] = 1
    return adj


def install_shims():
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")
        ts.scatter_sum = _scatter_sum
        ts.scatter_add = _scatter_sum
        ts.scatter_softmax = _scatter_softmax
        ts.scatter_max = _scatter_max
        ts.scatter_mean = _scatter_mean
        sys.modules["torch_scatter"] = ts
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tgu = types.ModuleType("torch_geometric.utils")
        tgu.to_dense_batch = _to_dense_batch
        tgu.to_dense_adj = _to_dense_adj
        tg.utils = tgu
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.utils"] = tgu


def load_reference(flavour="v1"):
    """Return the reference `models.*` modules (egnn, att_model, cross_att, model_utils) unmodified."""
    if not reference_available():
        raise RuntimeError("reference tree not present (only exists in the dev container)")
    install_shims()
    root = REF_V1 if flavour == "v1" else REF_PLUS
    # the two flavours share the package name `models`; purge before switching
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
    sys.path[:] = [p for p in sys.path if p not in (REF_V1, REF_PLUS)]
    sys.path.insert(0, root)
    mods = types.SimpleNamespace()
    for name in ("model_utils", "cross_att", "egnn", "att_model"):
        setattr(mods, name, importlib.import_module("models." + name))
    return mods


def _install_utils_stub():
    """`models/model.py:8` imports two helpers from utils.utils, whose own imports (rdkit, torchmetrics, ...) are
    not installed.  They are restated here from their definitions (utils/utils.py:147-158, 687-699)."""
    if "utils.utils" in sys.modules:
        return
    def get_keepNode_tensor(protein_node_xyz, pocket_radius, add_noise_to_com, chosen_pocket_com):
        if add_noise_to_com:
            chosen_pocket_com = chosen_pocket_com + add_noise_to_com * (2 * torch.rand_like(chosen_pocket_com) - 1)
        dis = torch.sqrt(torch.sum((protein_node_xyz - chosen_pocket_com.unsqueeze(0)) ** 2, dim=-1))
        return dis < pocket_radius

    def gumbel_softmax_no_random(logits, tau=1, hard=False, eps=1e-10, dim=-1):
        y_soft = (logits / tau).softmax(dim)
        if hard:
            index = y_soft.max(dim, keepdim=True)[1]
            y_hard = torch.zeros_like(logits, memory_format=torch.legacy_contiguous_format).scatter_(dim, index, 1.0)
            return y_hard - y_soft.detach() + y_soft
        return y_soft
    pkg = types.ModuleType("utils")
    mod = types.ModuleType("utils.utils")
    mod.get_keepNode_tensor = get_keepNode_tensor
    mod.gumbel_softmax_no_random = gumbel_softmax_no_random
    pkg.utils = mod
    sys.modules["utils"] = pkg
    sys.modules["utils.utils"] = mod


def load_reference_model_module(flavour="v1"):
    """The reference's L2 wrapper `models.model` (v1: IaBNet_..., plus: FABindPlus), imported unmodified."""
    mods = load_reference(flavour)
    if "mlflow" not in sys.modules:     # P/models/model.py:58-60 only calls mlflow.sklearn.autolog(disable=True); not installed here
        ml = types.ModuleType("mlflow")
        ml.sklearn = types.ModuleType("mlflow.sklearn")
        ml.sklearn.autolog = lambda **kw: None
        sys.modules["mlflow"] = ml
        sys.modules["mlflow.sklearn"] = ml.sklearn
    _install_utils_stub()
    mods.model = importlib.import_module("models.model")
    return mods


def published_args(**over):
    """The flags of the published v1 evaluation command (F/test_fabind.py:182) that the path reads."""
    from fabind_b200.config import published_args as _pa
    return _pa(**over)


def published_args_plus(**over):
    from fabind_b200.config import published_args_plus as _pa
    return _pa(**over)


def load_reference_post_optim():
    """`utils/post_optim_utils.py` of the v1 reference, imported unmodified (it imports rdkit only for its SDF writers)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    install_shims()
    for name in ("rdkit", "rdkit.Chem", "rdkit.Geometry"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["rdkit"].Chem = sys.modules["rdkit.Chem"]
    sys.modules["rdkit.Geometry"].Point3D = object
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_post_optim_utils", os.path.join(REF_V1, "utils", "post_optim_utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
