"""Pack a reference-layout state_dict into the flat weight arena of libfabind_b200.

The arena layout (slot names, shapes, offsets) is defined by the library (fb_weight_slot_*); this file
only says which reference parameters each slot is derived from.  Derivations (all exact in real
arithmetic, evaluated in float64 and rounded once):

 * first Linear of every per-edge MLP is split into per-node blocks + a rank-1 radial column
   (edge_mlp.0 at egnn.py:78, linear_kv at egnn.py:203-205 with its interleaved k/v rows);
 * the 32-channel InteractionModule output (cross_att.py:51) is folded into pair_transition.linear_1
   (K-concatenated operand, see csrc/layers.cu::pair_gather_kernel);
 * pair_transition.linear_2 followed by attn_bias_proj collapses to one vector (egnn.py:208,
   cross_att.py:53): bias = wb.(W2 t + b2) + bb = (W2^T wb).t + (wb.b2 + bb);
 * ac_u = coord_mlp.0.weight @ v_r (the radial column of v pushed through the next Linear).
"""
import ctypes as C

import torch

from . import _lib

HD = 128
QKX = 128   # extra columns of the stacked q|k GEMM (csrc/forward.cu)


def slots(hidden, n_layers, flavour=0, derived=False):
    """(name, rows, cols, offset) of the arena slots.  derived=False leaves out the library's own "f_*" slots (pre-multiplied
    projections filled on the device by fb_derive_weights, see `derive_on_device`): packers, their chain rules and the training path
    neither write nor read them."""
    l = _lib.lib()
    out = []
    name = C.create_string_buffer(64)
    r, c, o = C.c_int64(), C.c_int64(), C.c_int64()
    for i in range(l.fb_weight_slot_count_f(hidden, n_layers, flavour)):
        _lib.check(l.fb_weight_slot_info_f(hidden, n_layers, flavour, i, name, 64, C.byref(r), C.byref(c), C.byref(o)),
                   "fb_weight_slot_info_f")
        nm = name.value.decode()
        if not derived and nm.rpartition(".")[2].startswith("f_"):
            continue
        out.append((nm, r.value, c.value, o.value))
    return out


def base_elems(hidden, n_layers, flavour=0):
    """length of the base prefix of the arena: the library lays every derived ("f_*") slot out behind the last base slot"""
    return max(off + r * c for _, r, c, off in slots(hidden, n_layers, flavour))


def derive_on_device(w32, hidden, n_layers, flavour=0):
    """fill the derived ("f_*") slots of a device-resident fp32 arena in place (fb_derive_weights, on the current stream)"""
    if w32.device.type != "cuda" or w32.dtype != torch.float32 or not w32.is_contiguous():
        raise ValueError("derive_on_device: a contiguous float32 CUDA arena is required")
    st = C.c_void_p(torch.cuda.current_stream(w32.device).cuda_stream)
    with torch.cuda.device(w32.device):
        _lib.check(_lib.lib().fb_derive_weights(w32.data_ptr(), hidden, n_layers, flavour, st), "fb_derive_weights")
    return w32


def _gcl(sd, p, H):
    W1 = sd[p + "edge_mlp.0.weight"].double()
    return {
        "e1_rc": torch.cat([W1[:, :H], W1[:, H:2 * H]], 0), "e1_rad": W1[:, 2 * H], "e1_b": sd[p + "edge_mlp.0.bias"],
        "e2_w": sd[p + "edge_mlp.2.weight"], "e2_b": sd[p + "edge_mlp.2.bias"],
        "c1_w": sd[p + "coord_mlp.0.weight"], "c1_b": sd[p + "coord_mlp.0.bias"], "c2_w": sd[p + "coord_mlp.2.weight"][0],
        "n1_w": sd[p + "node_mlp.0.weight"], "n1_b": sd[p + "node_mlp.0.bias"],
        "n2_w": sd[p + "node_mlp.2.weight"], "n2_b": sd[p + "node_mlp.2.bias"],
    }


def _att(sd, p, H):
    ca = p + "cross_attn_module."
    pb, cb = ca + "p_attention_block.", ca + "c_attention_block."
    z = lambda n: torch.zeros(n, dtype=torch.float64)
    Wkv = sd[p + "linear_kv.weight"].double()
    bkv = sd[p + "linear_kv.bias"].double()
    W1p = sd[ca + "pair_transition.linear_1.weight"].double()
    W2 = sd[ca + "pair_transition.linear_2.weight"].double()
    b2 = sd[ca + "pair_transition.linear_2.bias"].double()
    wb = sd[p + "attn_bias_proj.weight"].double()[0]
    bb = sd[p + "attn_bias_proj.bias"].double()[0]
    ac1 = sd[p + "coord_mlp.0.weight"].double()
    v_r = Wkv[1::2, 0]
    return {
        "ca_c_w": torch.cat([sd[pb + "mha.linear_k.weight"], sd[pb + "mha.linear_v.weight"],
                             sd[cb + "mha.linear_q.weight"], sd[cb + "mha.linear_g.weight"]], 0),
        "ca_c_b": torch.cat([z(3 * HD), sd[cb + "mha.linear_g.bias"].double()]),
        "ca_p_w": torch.cat([sd[pb + "mha.linear_q.weight"], sd[pb + "mha.linear_g.weight"]], 0),
        "ca_p_b": torch.cat([z(HD), sd[pb + "mha.linear_g.bias"].double()]),
        "ca_p2_w": torch.cat([sd[cb + "mha.linear_k.weight"], sd[cb + "mha.linear_v.weight"]], 0),
        "o_p_w": sd[pb + "mha.linear_o.weight"], "o_p_b": sd[pb + "mha.linear_o.bias"],
        "o_c_w": sd[cb + "mha.linear_o.weight"], "o_c_b": sd[cb + "mha.linear_o.bias"],
        "tp1_w": sd[ca + "p_transition.linear_1.weight"], "tp1_b": sd[ca + "p_transition.linear_1.bias"],
        "tp2_w": sd[ca + "p_transition.linear_2.weight"], "tp2_b": sd[ca + "p_transition.linear_2.bias"],
        "tc1_w": sd[ca + "c_transition.linear_1.weight"], "tc1_b": sd[ca + "c_transition.linear_1.bias"],
        "tc2_w": sd[ca + "c_transition.linear_2.weight"], "tc2_b": sd[ca + "c_transition.linear_2.bias"],
        # first pair-transition Linear with the 32-channel interaction output folded in:
        #   W1 (pair0 + Wo t + bo) + b1 = [W1 | W1 Wo | 0] [pair0 | t | 0] + (b1 + W1 bo)
        "pt1_w": torch.cat([W1p, W1p @ sd[ca + "inter_layer.linear_out.weight"].double(), torch.zeros(2 * H, 32, dtype=torch.float64)], 1),
        "pt1_b": sd[ca + "pair_transition.linear_1.bias"].double() + W1p @ sd[ca + "inter_layer.linear_out.bias"].double(),
        "pt2v": W2.t() @ wb, "pt_c": (wb @ b2 + bb).reshape(1),
        # ONE stacked node GEMM:  q | k | inter_layer.linear_p (32) | inter_layer.linear_c (32) | zero pad  ||  v | vc
        # with vc = coord_mlp.0.weight @ (v rows): the Linear that follows v is folded (its bias is added per edge)
        "qk_w": torch.cat([sd[p + "linear_q.weight"].double(), Wkv[0::2, 1:],
                           sd[ca + "inter_layer.linear_p.weight"].double(), sd[ca + "inter_layer.linear_c.weight"].double(),
                           torch.zeros(QKX - 64, H, dtype=torch.float64), Wkv[1::2, 1:], ac1 @ Wkv[1::2, 1:]], 0),
        "qk_b": torch.cat([sd[p + "linear_q.bias"].double(), bkv[0::2], sd[ca + "inter_layer.linear_p.bias"].double(),
                           sd[ca + "inter_layer.linear_c.bias"].double(), z(QKX - 64), bkv[1::2], ac1 @ bkv[1::2]]),
        "k_r": Wkv[0::2, 0], "v_r": v_r,
        "ac1_b": sd[p + "coord_mlp.0.bias"], "ac2_w": sd[p + "coord_mlp.2.weight"][0],
        "ac_u": ac1 @ v_r,
    }


def _top(sd, H, L):
    rows, bias = [], []
    for l in range(L):
        ca = f"gnn.att_{l}.cross_attn_module."
        for blk in ("p_attention_block.", "c_attention_block."):
            rows += [sd[ca + blk + "linear.weight"], sd[ca + blk + "linear_g.weight"]]
            bias += [sd[ca + blk + "linear.bias"], sd[ca + blk + "linear_g.bias"]]
    d = {
        "in_w": sd["gnn.linear_in.weight"], "in_b": sd["gnn.linear_in.bias"],
        "out_w": sd["gnn.linear_out.weight"], "out_b": sd["gnn.linear_out.bias"],
        "il_p_w": sd["inter_layer.linear_p.weight"], "il_p_b": sd["inter_layer.linear_p.bias"],
        "il_c_w": sd["inter_layer.linear_c.weight"], "il_c_b": sd["inter_layer.linear_c.bias"],
        "il_o_w": sd["inter_layer.linear_out.weight"], "il_o_b": sd["inter_layer.linear_out.bias"],
    }
    if L > 0:
        pad = (-16 * L) % 128     # zero rows: the slot is padded to a multiple of 128 outputs
        d["pb_w"] = torch.cat(rows + [torch.zeros(pad, H)], 0)
        d["pb_b"] = torch.cat(bias + [torch.zeros(pad)], 0)
    return d


# ---- FABind+ layout (FABind_plus/fabind/models/model_utils.py:32-74: LayerNorm -> Linear -> ReLU -> Linear [-> ReLU]) ----
# A LayerNorm in front of a Linear is folded as  W LN(z) + b = rstd * ((W*gamma) z - mu * (W*gamma) 1) + (W beta + b);
# see csrc/plus.cu.  Explicit LayerNorms (coord_mlp / node_mlp of the GCL, the three transitions) keep gamma / beta.

def dp_of(H):
    """2H+1 edge-MLP features padded to a multiple of 64 (csrc/forward.cu::dp_of)"""
    return (2 * H + 1 + 63) // 64 * 64


def _pad_rows(t, rows):
    return torch.cat([t, torch.zeros((rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype)], 0)


def _gcl_plus(sd, p, H):
    D, Dp = 2 * H + 1, dp_of(H)
    g = sd[p + "edge_mlp.layernorm.weight"].double()
    bt = sd[p + "edge_mlp.layernorm.bias"].double()
    W1 = sd[p + "edge_mlp.linear1.weight"].double()
    if tuple(W1.shape) != (D, D):
        raise NotImplementedError("fabind_b200 (FABind+ layout): --mlp-hidden-scale 1 only (the published value)")
    W1g = W1 * g[None, :]
    W2 = sd[p + "edge_mlp.linear2.weight"].double()
    return {
        "e1_rc": torch.cat([_pad_rows(W1g[:, :H], Dp), _pad_rows(W1g[:, H:2 * H], Dp)], 0),
        "e1_rad": _pad_rows(W1g[:, 2 * H], Dp), "e1_g": _pad_rows(W1g.sum(1), Dp),
        "e1_c0": _pad_rows(W1 @ bt + sd[p + "edge_mlp.linear1.bias"].double(), Dp),
        "e2_w": torch.cat([W2, torch.zeros(H, Dp - D, dtype=torch.float64)], 1), "e2_b": sd[p + "edge_mlp.linear2.bias"],
        "cl_g": sd[p + "coord_mlp.layernorm.weight"], "cl_b": sd[p + "coord_mlp.layernorm.bias"],
        "c1_w": sd[p + "coord_mlp.linear1.weight"], "c1_b": sd[p + "coord_mlp.linear1.bias"],
        "c2_w": sd[p + "coord_mlp.linear2.weight"][0],
        "nl_g": sd[p + "node_mlp.layernorm.weight"], "nl_b": sd[p + "node_mlp.layernorm.bias"],
        "n1_w": sd[p + "node_mlp.linear1.weight"], "n1_b": sd[p + "node_mlp.linear1.bias"],
        "n2_w": sd[p + "node_mlp.linear2.weight"], "n2_b": sd[p + "node_mlp.linear2.bias"],
    }


def _att_plus(sd, p, H):
    ca = p + "cross_attn_module."
    pb, cb = ca + "p_attention_block.", ca + "c_attention_block."
    z = lambda n: torch.zeros(n, dtype=torch.float64)
    Wkv = sd[p + "linear_kv.weight"].double()
    bkv = sd[p + "linear_kv.bias"].double()
    v_r = Wkv[1::2, 0]
    ac1 = sd[p + "coord_mlp.linear1.weight"].double()
    ac1g = ac1 * sd[p + "coord_mlp.layernorm.weight"].double()[None, :]
    d = {
        "ca_c_w": torch.cat([sd[pb + "mha.linear_k.weight"], sd[pb + "mha.linear_v.weight"],
                             sd[cb + "mha.linear_q.weight"], sd[cb + "mha.linear_g.weight"]], 0),
        "ca_c_b": torch.cat([z(3 * HD), sd[cb + "mha.linear_g.bias"].double()]),
        "ca_p_w": torch.cat([sd[pb + "mha.linear_q.weight"], sd[pb + "mha.linear_g.weight"]], 0),
        "ca_p_b": torch.cat([z(HD), sd[pb + "mha.linear_g.bias"].double()]),
        "ca_p2_w": torch.cat([sd[cb + "mha.linear_k.weight"], sd[cb + "mha.linear_v.weight"]], 0),
        "o_p_w": sd[pb + "mha.linear_o.weight"], "o_p_b": sd[pb + "mha.linear_o.bias"],
        "o_c_w": sd[cb + "mha.linear_o.weight"], "o_c_b": sd[cb + "mha.linear_o.bias"],
        # gated pair-bias projections of THIS layer's two RowAttentionBlocks (the pair embedding changes per layer)
        "pb_w": _pad_rows(torch.cat([sd[pb + "linear.weight"], sd[pb + "linear_g.weight"],
                                     sd[cb + "linear.weight"], sd[cb + "linear_g.weight"]], 0), 128),
        "pb_b": _pad_rows(torch.cat([sd[pb + "linear.bias"], sd[pb + "linear_g.bias"],
                                     sd[cb + "linear.bias"], sd[cb + "linear_g.bias"]], 0), 128),
        "zo_w": sd[ca + "inter_layer.linear_out.weight"].t().contiguous(), "zo_b": sd[ca + "inter_layer.linear_out.bias"],
        "zl_g": sd[ca + "pair_transition.layernorm.weight"], "zl_b": sd[ca + "pair_transition.layernorm.bias"],
        "pt1_w": sd[ca + "pair_transition.linear1.weight"], "pt1_b": sd[ca + "pair_transition.linear1.bias"],
        "pt2_w": sd[ca + "pair_transition.linear2.weight"], "pt2_b": sd[ca + "pair_transition.linear2.bias"],
        "wb": sd[p + "attn_bias_proj.weight"][0], "pt_c": sd[p + "attn_bias_proj.bias"].reshape(1),
        # q | k | inter_layer.linear_p (32) | inter_layer.linear_c (32) | zero pad  ||  v | vc, vc = (linear1*gamma) v
        "qk_w": torch.cat([sd[p + "linear_q.weight"].double(), Wkv[0::2, 1:],
                           sd[ca + "inter_layer.linear_p.weight"].double(), sd[ca + "inter_layer.linear_c.weight"].double(),
                           torch.zeros(QKX - 64, H, dtype=torch.float64), Wkv[1::2, 1:], ac1g @ Wkv[1::2, 1:]], 0),
        "qk_b": torch.cat([sd[p + "linear_q.bias"].double(), bkv[0::2], sd[ca + "inter_layer.linear_p.bias"].double(),
                           sd[ca + "inter_layer.linear_c.bias"].double(), z(QKX - 64), bkv[1::2], ac1g @ bkv[1::2]]),
        "k_r": Wkv[0::2, 0], "v_r": v_r,
        "ac_c0": ac1 @ sd[p + "coord_mlp.layernorm.bias"].double() + sd[p + "coord_mlp.linear1.bias"].double(),
        "ac2_w": sd[p + "coord_mlp.linear2.weight"][0], "ac_u": ac1g @ v_r, "ac_g": ac1g.sum(1),
        "ac_r": torch.stack([v_r.sum(), (v_r * v_r).sum()]),
    }
    for t, key in (("tp", "p_transition."), ("tc", "c_transition.")):
        d[t + "l_g"], d[t + "l_b"] = sd[ca + key + "layernorm.weight"], sd[ca + key + "layernorm.bias"]
        d[t + "1_w"], d[t + "1_b"] = sd[ca + key + "linear1.weight"], sd[ca + key + "linear1.bias"]
        d[t + "2_w"], d[t + "2_b"] = sd[ca + key + "linear2.weight"], sd[ca + key + "linear2.bias"]
    return d


def pack_state_dict(sd, hidden, n_layers, flavour=0, differentiable=False, device="cpu"):
    """Returns the fp32 arena (CPU tensor) for a state_dict with the reference's key names
    (flavour 0: FABind v1 layout, 1: FABind+ layout).  differentiable=True keeps the autograd graph from the reference
    parameters to the arena (every derivation is a torch expression): gradients w.r.t. arena slots map back to reference
    parameters by the chain rule -- how a training path in this formulation returns `state_dict`-shaped gradients
    (tests/test_formulation_cpu.py::test_refactored_formulation_gradients)."""
    dev = torch.device(device)
    sd = {k: (v.to(dev) if differentiable else v.detach().to(dev)) for k, v in sd.items()}
    l = _lib.lib()
    with torch.device(dev):      # every factory call of the derivations below lands on `device` (training re-packs on the GPU every step)
        return _pack_on_device(sd, l, hidden, n_layers, flavour)


def _pack_on_device(sd, l, hidden, n_layers, flavour):
    arena = torch.zeros(l.fb_weight_arena_elems_f(hidden, n_layers, flavour), dtype=torch.float32)
    gcl, att = (_gcl_plus, _att_plus) if flavour == 1 else (_gcl, _att)
    groups = {"": _top(sd, hidden, 0 if flavour == 1 else n_layers)}
    for i in range(n_layers):
        groups[f"gcl{i}."] = gcl(sd, f"gnn.gcl_{i}.", hidden)
        groups[f"att{i}."] = att(sd, f"gnn.att_{i}.", hidden)
    groups["out."] = gcl(sd, "gnn.out_layer.", hidden)
    for name, rows, cols, off in slots(hidden, n_layers, flavour):
        pre, _, base = name.rpartition(".")
        pre = pre + "." if pre else ""
        if rows * cols == 0:
            continue
        t = groups[pre][base]
        if t.numel() != rows * cols or (t.dim() == 2 and tuple(t.shape) != (rows, cols)):
            raise RuntimeError(f"weight slot {name}: library expects [{rows},{cols}], packer produced {tuple(t.shape)}")
        arena[off:off + rows * cols] = t.reshape(-1).to(torch.float32)
    return arena


def arena_grads_to_state_dict(sd, arena_grad, hidden, n_layers, flavour=0, device="cpu"):
    """Chain rule of the packer: gradient w.r.t. the flat weight arena (what backward kernels of this formulation produce:
    hoisted first Linears, folded LayerNorms, collapsed pair-bias vector, stacked projections) -> gradients keyed and shaped
    like the reference `state_dict` (what an optimizer over the drop-in modules' parameters consumes).  Parameters the
    formulation never reads (`att_i.inter_layer.*` of the v1 layout, unused in the reference too) come back as zeros, which is
    what the flat gradient all-reduce (`shard.allreduce_gradients`) expects.  float32 like the arena.  On the CPU this costs seconds at
    the published size (float64 derivations, 28 M arena elements); the training step runs it on the GPU (`device=`)."""
    dev = torch.device(device)
    leaves = {k: v.detach().to(dev, torch.float32).clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    full = dict(sd)
    full.update(leaves)
    with torch.enable_grad():     # also callable from inside an autograd.Function's backward (grad mode is off there)
        arena = pack_state_dict(full, hidden, n_layers, flavour, differentiable=True, device=dev)
        arena.backward(arena_grad.detach().to(dev, torch.float32).reshape(-1))
    return {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}


class GraphedPacker:
    """`pack_state_dict` and `arena_grads_to_state_dict` of ONE module, captured once as CUDA graphs and replayed every training step.

    Why: a drop-in for the reference's training scripts keeps `state_dict`-shaped parameters (torch optimizers step them), so every
    step packs them into the arena and pushes the arena gradient back through the packer's chain rule.  Both are a few hundred small
    torch ops (6 ms + 22 ms of host time at the published size, the GPU mostly idle); replayed from a graph they cost the sum of their
    kernels.  The graphs read the LIVE parameter storages (in-place optimizer updates are seen; a re-allocated parameter changes
    `key` and triggers a re-capture) and write static buffers; callers get fresh clones, so nothing aliases across steps.
    Falls back to the eager functions when capture is not possible (CPU tensors, capture error)."""

    def __init__(self, sd, hidden, n_layers, flavour, device):
        self.hidden, self.n_layers, self.flavour, self.device = hidden, n_layers, flavour, torch.device(device)
        self.sd = dict(sd)
        self.names = [k for k, v in sd.items() if v.is_floating_point()]
        self.key = self.make_key(sd, hidden, n_layers, flavour, device)
        self.layout, o = [], 0
        for k in self.names:
            self.layout.append((k, o, sd[k].numel(), tuple(sd[k].shape)))
            o += sd[k].numel()
        self.total = o
        self.ok = False
        self.error = None
        if self.device.type == "cuda" and all(v.device == self.device for v in sd.values()) and \
                all(sd[k].dtype == torch.float32 for k in self.names):
            try:
                self._capture()
                self.ok = True
            except Exception as e:          # capture is an optimisation: the eager path below is always correct
                self.error = repr(e)
                torch.cuda.synchronize(self.device)

    @staticmethod
    def make_key(sd, hidden, n_layers, flavour, device):
        return (hidden, n_layers, flavour, str(device), tuple((k, v.data_ptr(), tuple(v.shape), v.dtype) for k, v in sd.items()))

    def _eager_unpack_flat(self, garena):
        g = arena_grads_to_state_dict(self.sd, garena, self.hidden, self.n_layers, self.flavour, device=self.device)
        return torch.cat([g[k].reshape(-1) for k in self.names])

    def _capture(self):
        dev = self.device
        n_arena = _lib.lib().fb_weight_arena_elems_f(self.hidden, self.n_layers, self.flavour)
        self.garena = torch.zeros(n_arena, dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):                                   # warm-up outside the capture (cuBLAS handles, autograd threads)
                pack_state_dict(self.sd, self.hidden, self.n_layers, self.flavour, device=dev)
                self._eager_unpack_flat(self.garena)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.g_pack = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_pack):
            self.arena_static = pack_state_dict(self.sd, self.hidden, self.n_layers, self.flavour, device=dev)
        self.g_unpack = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_unpack):
            self.flat_static = self._eager_unpack_flat(self.garena)

    def pack(self):
        if not self.ok:
            return pack_state_dict(self.sd, self.hidden, self.n_layers, self.flavour, device=self.device)
        self.g_pack.replay()
        return self.arena_static.clone()

    def unpack(self, garena):
        """arena gradient -> (flat fp32 gradient of all floating parameters in state_dict order, {name: view into it})"""
        if self.ok:
            self.garena.copy_(garena.reshape(-1))
            self.g_unpack.replay()
            flat = self.flat_static.clone()
        else:
            flat = self._eager_unpack_flat(garena)
        return flat, {k: flat[o:o + n].view(shape) for k, o, n, shape in self.layout}


class FastPackerV1:
    """Explicit pack / chain rule for the v1 layout (training path): the packer is a SELECTION of parameter elements (with zero pads)
    plus seven small bilinear derivations per MC_Att_L.  The selection is found once by running the derivations above on a
    state_dict whose entries hold their own flat index (float64 ids), so it stays tied to `_top/_gcl/_att`; per step
        pack:    arena[dst] = theta[src]  (one gather)          + the derived blocks (a few fp64 matmuls per layer)
        unpack:  g_theta.index_add_(src, g_arena[dst])          + their hand-written chain rule
    instead of several thousand small autograd kernels (`arena_grads_to_state_dict`: 14-18 ms of GPU time at the published size even
    when replayed from a CUDA graph).  Checked against the generic functions in tests/test_host_logic.py."""

    def __init__(self, sd, hidden, n_layers, device):
        H, L = hidden, n_layers
        self.H, self.L, self.device = H, L, torch.device(device)
        self.sd = dict(sd)
        self.ok, self.error = True, None
        self.names = [k for k, v in sd.items() if v.is_floating_point()]
        self.key = GraphedPacker.make_key(sd, H, L, 0, device)
        self.layout, o = [], 0
        ids = {}
        for k in self.names:
            n = sd[k].numel()
            self.layout.append((k, o, n, tuple(sd[k].shape)))
            ids[k] = (torch.arange(n, dtype=torch.float64) + (o + 1)).view(sd[k].shape)
            o += n
        self.total = o
        self.off = {k: (o_, n_, s_) for k, o_, n_, s_ in self.layout}
        groups = {"": _top(ids, H, L)}
        for i in range(L):
            groups[f"gcl{i}."] = _gcl(ids, f"gnn.gcl_{i}.", H)
            groups[f"att{i}."] = _att(ids, f"gnn.att_{i}.", H)
        groups["out."] = _gcl(ids, "gnn.out_layer.", H)
        self.n_arena = _lib.lib().fb_weight_arena_elems_f(H, L, 0)
        src = torch.zeros(self.n_arena, dtype=torch.int64)
        self.slot = {}
        for name, rows, cols, off in slots(H, L, 0):
            if rows * cols == 0:
                continue
            pre, _, base = name.rpartition(".")
            pre = pre + "." if pre else ""
            self.slot[name] = (off, rows, cols)
            blk = groups[pre][base].reshape(-1).double().round().long()
            if pre.startswith("att"):
                v = blk.view(rows, cols)
                if base == "pt1_w":
                    v[:, H:] = 0
                elif base in ("pt1_b", "pt2v", "pt_c", "ac_u"):
                    v[:] = 0
                elif base == "qk_w":
                    v[3 * H + QKX:] = 0
                elif base == "qk_b":
                    v[:, 3 * H + QKX:] = 0
            src[off:off + rows * cols] = blk.clamp(min=0, max=self.total)
        dst = src.nonzero().squeeze(1)
        self.dst = dst.to(self.device)
        self.src = (src[dst] - 1).to(self.device)

    # ---- the seven derived blocks of one MC_Att_L (weights.py::_att) and their chain rule -------------------------------------
    def _att_keys(self, l):
        p = f"gnn.att_{l}."
        ca = p + "cross_attn_module."
        return dict(W1p=ca + "pair_transition.linear_1.weight", b1=ca + "pair_transition.linear_1.bias",
                    W2=ca + "pair_transition.linear_2.weight", b2=ca + "pair_transition.linear_2.bias",
                    Wo=ca + "inter_layer.linear_out.weight", bo=ca + "inter_layer.linear_out.bias",
                    wb=p + "attn_bias_proj.weight", bb=p + "attn_bias_proj.bias", ac1=p + "coord_mlp.0.weight",
                    Wkv=p + "linear_kv.weight", bkv=p + "linear_kv.bias")

    def _view(self, flat, key):
        o, n, shape = self.off[key]
        return flat[o:o + n].view(shape)

    def _slot(self, arena, name):
        off, r, c = self.slot[name]
        return arena[off:off + r * c].view(r, c)

    # ---- all L layers at once: the parameters of att_0 .. att_{L-1} sit at a constant stride in the flat state_dict order, and so do
    # their arena slots, so [L, ...] views exist without a copy (checked once; otherwise the per-layer loops below run)
    def _layers_stackable(self):
        ok = getattr(self, "_stack_ok", None)
        if ok is not None:
            return ok
        ok = self.L >= 1
        self._pstride, self._sstride = 0, 0
        if self.L > 1:
            keys = [self._att_keys(l) for l in range(self.L)]
            d = {self.off[keys[l + 1][n]][0] - self.off[keys[l][n]][0] for l in range(self.L - 1) for n in keys[0]}
            ds = {self.slot[f"att{l + 1}.{b}"][0] - self.slot[f"att{l}.{b}"][0] for l in range(self.L - 1)
                  for b in ("pt1_w", "pt1_b", "pt2v", "pt_c", "qk_w", "qk_b", "ac_u")}
            ok = len(d) == 1 and len(ds) == 1
            if ok:
                self._pstride, self._sstride = d.pop(), ds.pop()
        self._stack_ok = ok
        return ok

    @staticmethod
    def _cstrides(shape):
        st, acc = [], 1
        for n in reversed(shape):
            st.append(acc)
            acc *= n
        return tuple(reversed(st))

    def _pstack(self, flat, n):
        """[L, *shape] view of parameter `n` (short name of _att_keys) of all layers inside a flat state_dict-ordered vector"""
        o, _, shape = self.off[self._att_keys(0)[n]]
        return flat.as_strided((self.L,) + tuple(shape), (self._pstride,) + self._cstrides(shape), flat.storage_offset() + o)

    def _sstack(self, arena, base):
        """[L, rows, cols] view of slot att<l>.<base> of all layers inside a flat arena"""
        off, r, c = self.slot["att0." + base]
        return arena.as_strided((self.L, r, c), (self._sstride, c, 1), arena.storage_offset() + off)

    def _theta_from(self, sd):
        """flat fp32 copy of the floating parameters in state_dict order, in ONE persistent buffer (a few fused multi-tensor copies
        instead of one conversion + one slice of a concatenation per parameter)"""
        th = getattr(self, "_theta", None)
        if th is None:
            th = self._theta = torch.empty(self.total, dtype=torch.float32, device=self.device)
            self._theta_views = [th[o:o + n].view(shape) for _, o, n, shape in self.layout]
        src = [sd[k].detach() for k in self.names]
        if all(t.device == self.device and t.dtype == torch.float32 for t in src):
            torch._foreach_copy_(self._theta_views, src)
        else:
            for v, t in zip(self._theta_views, src):
                v.copy_(t)
        return th

    def pack(self, sd=None):
        """fp32 arena on self.device from the live state_dict tensors (same values as pack_state_dict)"""
        H = self.H
        sd = self.sd if sd is None else sd
        theta = self._theta_from(sd)
        # ONE arena buffer per packer, rewritten in place every step: every element that is not a zero pad is overwritten below, the
        # pads are zeroed once.  (Two forwards before a backward see the same weights, hence the same values; train.slot_tensors keys
        # its cached views on this buffer.)
        arena = getattr(self, "_arena", None)
        if arena is None:
            arena = self._arena = torch.zeros(self.n_arena, dtype=torch.float32, device=self.device)
        arena[self.dst] = theta[self.src]
        if self._layers_stackable():
            t = {n: self._pstack(theta, n).double() for n in self._att_keys(0)}
            Wkv, bkv, ac1 = t["Wkv"], t["bkv"], t["ac1"]
            wb = t["wb"][:, 0]                                                        # [L, 2H]
            S = lambda base: self._sstack(arena, base)
            S("pt1_w")[:, :, H:H + 32] = (t["W1p"] @ t["Wo"]).float()
            S("pt1_b")[:, 0] = (t["b1"] + (t["W1p"] @ t["bo"].unsqueeze(-1)).squeeze(-1)).float()
            S("pt2v")[:, 0] = (t["W2"].transpose(1, 2) @ wb.unsqueeze(-1)).squeeze(-1).float()
            S("pt_c")[:, 0, 0] = ((wb * t["b2"]).sum(-1) + t["bb"][:, 0]).float()
            S("qk_w")[:, 3 * H + QKX:] = (ac1 @ Wkv[:, 1::2, 1:]).float()
            S("qk_b")[:, 0, 3 * H + QKX:] = (ac1 @ bkv[:, 1::2].unsqueeze(-1)).squeeze(-1).float()
            S("ac_u")[:, 0] = (ac1 @ Wkv[:, 1::2, 0].unsqueeze(-1)).squeeze(-1).float()
            return arena
        for l in range(self.L):
            k = self._att_keys(l)
            t = {n: self._view(theta, key).double() for n, key in k.items()}
            pre = f"att{l}."
            Wkv, bkv, ac1 = t["Wkv"], t["bkv"], t["ac1"]
            wb = t["wb"][0]
            self._slot(arena, pre + "pt1_w")[:, H:H + 32] = (t["W1p"] @ t["Wo"]).float()
            self._slot(arena, pre + "pt1_b")[0] = (t["b1"] + t["W1p"] @ t["bo"]).float()
            self._slot(arena, pre + "pt2v")[0] = (t["W2"].t() @ wb).float()
            self._slot(arena, pre + "pt_c")[0, 0] = (wb @ t["b2"] + t["bb"][0]).float()
            self._slot(arena, pre + "qk_w")[3 * H + QKX:] = (ac1 @ Wkv[1::2, 1:]).float()
            self._slot(arena, pre + "qk_b")[0, 3 * H + QKX:] = (ac1 @ bkv[1::2]).float()
            self._slot(arena, pre + "ac_u")[0] = (ac1 @ Wkv[1::2, 0]).float()
        return arena

    def unpack(self, garena, sd=None):
        """arena gradient -> (flat fp32 gradient in state_dict order, {name: view})"""
        H = self.H
        sd = self.sd if sd is None else sd
        garena = garena.reshape(-1).to(self.device, torch.float32)
        g = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        g.index_add_(0, self.src, garena[self.dst])
        if self._layers_stackable() and garena.is_contiguous():
            theta = self._theta_from(sd)
            P = {n: self._pstack(theta, n) for n in self._att_keys(0)}
            G = {n: self._pstack(g, n) for n in self._att_keys(0)}
            S = lambda base: self._sstack(garena, base)
            col = lambda v: v.unsqueeze(-1)                                           # [L, n] -> [L, n, 1]
            outer = lambda a, b: a.unsqueeze(-1) * b.unsqueeze(-2)                     # batched torch.outer
            wb = P["wb"][:, 0]
            g_pt1 = S("pt1_w")[:, :, H:H + 32]                                         # d(W1p Wo)
            G["W1p"] += g_pt1 @ P["Wo"].transpose(1, 2)
            G["Wo"] += P["W1p"].transpose(1, 2) @ g_pt1
            g_b = S("pt1_b")[:, 0]                                                     # d(b1 + W1p bo)
            G["b1"] += g_b
            G["W1p"] += outer(g_b, P["bo"])
            G["bo"] += (P["W1p"].transpose(1, 2) @ col(g_b)).squeeze(-1)
            g_v = S("pt2v")[:, 0]                                                      # d(W2^T wb)
            G["W2"] += outer(wb, g_v)
            G["wb"][:, 0] += (P["W2"] @ col(g_v)).squeeze(-1)
            g_c = S("pt_c")[:, 0, 0]                                                   # d(wb . b2 + bb)
            G["wb"][:, 0] += col(g_c) * P["b2"]
            G["b2"] += col(g_c) * wb
            G["bb"][:, 0] += g_c
            g_vc = S("qk_w")[:, 3 * H + QKX:]                                          # d(ac1 Wkv_v)
            G["ac1"] += g_vc @ P["Wkv"][:, 1::2, 1:].transpose(1, 2)
            G["Wkv"][:, 1::2, 1:] += P["ac1"].transpose(1, 2) @ g_vc
            g_vb = S("qk_b")[:, 0, 3 * H + QKX:]                                       # d(ac1 bkv_v)
            G["ac1"] += outer(g_vb, P["bkv"][:, 1::2])
            G["bkv"][:, 1::2] += (P["ac1"].transpose(1, 2) @ col(g_vb)).squeeze(-1)
            g_u = S("ac_u")[:, 0]                                                      # d(ac1 v_r)
            G["ac1"] += outer(g_u, P["Wkv"][:, 1::2, 0])
            G["Wkv"][:, 1::2, 0] += (P["ac1"].transpose(1, 2) @ col(g_u)).squeeze(-1)
            return g, {k_: g[o:o + n].view(shape) for k_, o, n, shape in self.layout}
        for l in range(self.L):
            k = self._att_keys(l)
            P = {n: sd[key].detach().to(self.device, torch.float32) for n, key in k.items()}
            G = {n: self._view(g, key) for n, key in k.items()}
            pre = f"att{l}."
            wb = P["wb"][0]
            g_pt1 = self._slot(garena, pre + "pt1_w")[:, H:H + 32]           # d(W1p Wo)
            G["W1p"] += g_pt1 @ P["Wo"].t()
            G["Wo"] += P["W1p"].t() @ g_pt1
            g_b = self._slot(garena, pre + "pt1_b")[0]                        # d(b1 + W1p bo)
            G["b1"] += g_b
            G["W1p"] += torch.outer(g_b, P["bo"])
            G["bo"] += P["W1p"].t() @ g_b
            g_v = self._slot(garena, pre + "pt2v")[0]                         # d(W2^T wb)
            G["W2"] += torch.outer(wb, g_v)
            G["wb"][0] += P["W2"] @ g_v
            g_c = self._slot(garena, pre + "pt_c")[0, 0]                      # d(wb . b2 + bb)
            G["wb"][0] += g_c * P["b2"]
            G["b2"] += g_c * wb
            G["bb"][0] += g_c
            g_vc = self._slot(garena, pre + "qk_w")[3 * H + QKX:]             # d(ac1 Wkv_v)
            G["ac1"] += g_vc @ P["Wkv"][1::2, 1:].t()
            G["Wkv"][1::2, 1:] += P["ac1"].t() @ g_vc
            g_vb = self._slot(garena, pre + "qk_b")[0, 3 * H + QKX:]          # d(ac1 bkv_v)
            G["ac1"] += torch.outer(g_vb, P["bkv"][1::2])
            G["bkv"][1::2] += P["ac1"].t() @ g_vb
            g_u = self._slot(garena, pre + "ac_u")[0]                         # d(ac1 v_r)
            G["ac1"] += torch.outer(g_u, P["Wkv"][1::2, 0])
            G["Wkv"][1::2, 0] += P["ac1"].t() @ g_u
        return g, {k_: g[o:o + n].view(shape) for k_, o, n, shape in self.layout}
