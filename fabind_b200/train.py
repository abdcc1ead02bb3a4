"""Assembly of one training step through the docking stack (v1 and FABind+ layouts) (BASELINE config 5: forward + backward + gradient all-reduce).

Reference semantics (att_model.py:210-246, refine='refine_coord'): the first `n_iter - 1` refinement iterations run under no_grad
and only move the ligand; the last one is differentiated.  So a step is
  1. iterations 0 .. n_iter-2: the inference path (one fb_model_forward call, training-mode dropout masks included);
  2. edge lists of the last iteration from the graph builder (`ComplexGraph.construct_edges`, bit-exact), mapped to the internal
     node order;
  3. `backward.stack_forward_train_v1` (training-mode forward of the last iteration, keeps what the reverse pass needs);
  4. the caller's loss on (X, H) -> gradients of the outputs (the losses of main_fabind.py:380-401 live outside the path);
  5. `backward.stack_backward_v1` -> gradient of the weight arena -> `weights.arena_grads_to_state_dict` -> parameter `.grad`s;
  6. `shard.allreduce_gradients` over the ranks (one flat NCCL collective).
`n_iter`: the reference draws randint(1, n_iter) per step when --random-n-iter is set (att_model.py:210-211); `forward_with_grad`
does the same.  Dropout (v1 layout): the reference trains the stack with nn.Dropout(0.1) at egnn.py:82,106,236,398,461 and
cross_att.py:128; the same sites carry the library's counter-based masks here, in the no_grad iterations (fused into the inference
kernels' epilogues), in the training-mode forward and -- the identical mask function -- in the reverse pass.  The FABind+ layout's
reverse pass carries no masks yet: `plus.EfficientMCAttModel` refuses train()+autograd with dropout_p > 0.

Status: parity-green on a B200 (tests/test_gpu_train_forward.py: training-mode forward, reverse pass and the assembled step against
the unmodified reference's parameter gradients, both layouts) and on the CPU with the kernel wrappers replaced by their
specifications (tests/test_backward_orchestration.py).  `EfficientMCAttModel.forward` routes here in train() mode with autograd.
"""
import torch

from . import backward as bw
from .layout import build_layout
from .weights import slots, base_elems, pack_state_dict, arena_grads_to_state_dict, GraphedPacker, FastPackerV1

# per-step pack / un-pack of the weight arena: v1 layout = explicit selection + hand-written chain rule (weights.FastPackerV1), FABind+
# layout = the generic functions replayed from CUDA graphs (weights.GraphedPacker); False = the generic eager torch ops every step
USE_FAST_PACKER = True


def _packer(model, sd, H, L, flavour, dev):
    """the module's packer object, rebuilt when a parameter storage was re-allocated"""
    key = GraphedPacker.make_key(sd, H, L, flavour, dev)
    pk = getattr(model, "_fb_packer", None)
    if pk is None or pk.key != key:
        if flavour == 0 and all(v.dtype == torch.float32 for v in sd.values() if v.is_floating_point()):
            pk = FastPackerV1(sd, H, L, dev)
        else:
            pk = GraphedPacker(sd, H, L, flavour, dev)
        try:
            object.__setattr__(model, "_fb_packer", pk)
        except Exception:
            pass
    return pk


_TDESC = {}


def _transposed_twins(arena, hidden, n_layers, flavour, out=None):
    """bf16 buffer holding, at every matrix slot's own offset, the slot TRANSPOSED ([cols, rows]): one launch for the whole arena
    (fb_transpose_slots_bf16) instead of one transpose-copy per slot (~80 launches and 1.5 ms per step at the published size)"""
    from . import _lib
    key = (hidden, n_layers, flavour, arena.device)
    d = _TDESC.get(key)
    if d is None:
        rows, tb = [], [0]
        for name, r, c, off in slots(hidden, n_layers, flavour):
            if r > 1 and r * c > 0:
                rows.append((off, r, c, off))
                tb.append(tb[-1] + ((r + 31) // 32) * ((c + 31) // 32))
        d = _TDESC[key] = (torch.tensor(rows, dtype=torch.int64).reshape(-1).to(arena.device), torch.tensor(tb, dtype=torch.int32).to(arena.device),
                           len(rows), tb[-1])
    desc, tbeg, n, n_tiles = d
    if out is None:
        out = torch.empty(base_elems(hidden, n_layers, flavour), dtype=torch.bfloat16, device=arena.device)
    _lib.check(_lib.lib().fb_transpose_slots_bf16(arena.data_ptr(), desc.data_ptr(), tbeg.data_ptr(), n, n_tiles, out.data_ptr(),
                                                  bw._st(arena)), "fb_transpose_slots_bf16")
    return out


_SLOT_CACHE = {}


def slot_tensors(arena, hidden, n_layers, flavour=0, bf16=None):
    """{prefix: {slot: tensor view (+ `_t` transposed copies of the matrices)}} over a flat arena on any device.
    bf16 (default: backward.PRECISION == "bf16"): the arena is converted to bf16 ONCE, every matrix view gets its bf16 twin registered
    with backward.register_bf16 (the GEMM wrappers pick it up instead of converting per call) and the `_t` transposes are made in bf16
    only (their one consumer is the data-gradient GEMM), all of them by one launch (fb_transpose_slots_bf16).
    A packer that rewrites ONE arena buffer in place (weights.FastPackerV1) gets the dictionary of views back from a cache keyed on
    that buffer: per step only the two conversion launches run (the Python loop over 184 slots cost 2.2 ms per step)."""
    bf16 = (bw.PRECISION == "bf16") if bf16 is None else bf16
    bf16 = bf16 and arena.is_cuda
    key = (arena.data_ptr(), arena.numel(), arena.device, hidden, n_layers, flavour, bf16)
    hit = _SLOT_CACHE.get(key) if bf16 else None
    if hit is not None and hit["arena"] is arena:
        nb = base_elems(hidden, n_layers, flavour)
        hit["a16"].copy_(arena[:nb])
        _transposed_twins(arena, hidden, n_layers, flavour, out=hit["t16"])
        if bw._W16 is not hit["reg"]:               # another arena was laid out in between: bring this one's twins back
            bw._W16 = hit["reg"]
        return hit["out"]
    out = {}
    bw._W16 = {}            # a fresh registry (the cached entry of another arena keeps its own dictionary)
    # (only the base prefix: the derived slots behind it belong to the inference path, fb_derive_weights)
    a16 = arena[:base_elems(hidden, n_layers, flavour)].to(torch.bfloat16) if bf16 else None
    t16 = _transposed_twins(arena, hidden, n_layers, flavour) if bf16 else None
    for name, r, c, off in slots(hidden, n_layers, flavour):
        if r * c == 0:
            continue
        pre, _, base = name.rpartition(".")
        pre = pre + "." if pre else ""
        t = arena[off:off + r * c]
        t = t.view(r, c) if r > 1 else t.view(c)
        d = out.setdefault(pre, {})
        d[base] = t.contiguous()
        if r > 1:
            if bf16:
                bw.register_bf16(d[base], a16[off:off + r * c].view(r, c))
                d[base + "_t"] = t16[off:off + r * c].view(c, r)
            else:
                d[base + "_t"] = t.t().contiguous()
    if bf16:
        _SLOT_CACHE.clear()                          # one live arena at a time is the training loop's pattern
        _SLOT_CACHE[key] = dict(arena=arena, a16=a16, t16=t16, out=out, reg=bw._W16)
    return out


def internal_graph(lay, ctx_edges, inter_edges, bonds, las, device):
    """reference-order edge lists in caller node ids (construct_edges + the bond list the caller prepends, att_model.py:231) ->
    int32 lists in internal ids; the interface edges sorted by destination row (segment softmax wants CSR order)"""
    o = lay.offs
    N, B = lay.N, lay.B
    blob = lay.blob.to(device)
    inv = blob[o["inv"]:o["inv"] + N].long()
    i32 = lambda t: t.to(torch.int32).contiguous()
    ctx = torch.cat([bonds.to(device), ctx_edges.to(device)], dim=1)
    ctx_row, ctx_col = inv[ctx[0]], inv[ctx[1]]
    ir, ic = inv[inter_edges[0].to(device)], inv[inter_edges[1].to(device)]
    order = torch.sort(ir, stable=True).indices
    las_a, las_b = inv[las[0].to(device)], inv[las[1].to(device)]
    edges = dict(ctx_row=i32(ctx_row), ctx_col=i32(ctx_col), int_row=i32(ir[order]), int_col=i32(ic[order]), las_a=i32(las_a), las_b=i32(las_b))
    geo = dict(Nc=lay.Nc_tot, B=B, max_c=lay.max_c, max_p=lay.max_p,
               node_cplx=blob[o["node_cplx"]:o["node_cplx"] + N].contiguous(), c_off=blob[o["c_off"]:o["c_off"] + B + 1].contiguous(),
               p_off=blob[o["p_off"]:o["p_off"] + B + 1].contiguous(), pair_base=blob[o["pair_base"]:o["pair_base"] + B + 1].contiguous())
    perm = blob[o["perm"]:o["perm"] + N].long()
    moves = (lay.flags.to(device) & 4) != 0
    return geo, edges, perm, moves


def _gpu_prev_coords(model, fa, n_iter=None, dropout=None):
    """iterations 0 .. n_iter-2 through the inference path (in place on a copy of X; the module's mode flags and configuration are
    not touched).  dropout = (p, seed, colonly) or None: training-mode masks of those iterations."""
    from .runtime import model_forward
    n = model._cfg["n_iter"] if n_iter is None else int(n_iter)
    X = fa["X"].detach().clone()
    if n <= 1:
        return X
    plus = int(model._cfg.get("flavour", 0)) == 1
    with torch.no_grad():
        model_forward(model, model._packed, X, fa["H"].detach(), fa["batch_id"], fa["segment_id"], fa["mask"], fa["is_global"],
                      fa["compound_edge_index"], fa["LAS_edge_index"], fa["batched_complex_coord_LAS"], model._cfg,
                      getattr(model, "precision", "fp32"), want_pair=False if plus else None, dropout=dropout, n_iter=n - 1)
    return X


def _gpu_edges(model, X_prev, fa):
    ctx, inter, _ = model.extract_edges.construct_edges(X_prev, fa["batch_id"], fa["segment_id"], fa["is_global"])
    return ctx, inter


def _forward_half(model, fa, prev_coords=None, edge_lists=None, state_dict=None, n_iter=None, dropout=None):
    """steps 1-3: everything up to the outputs; returns (X_out, H_out, pair rows or None, state for the reverse half).
    prev_coords(model, fa) / edge_lists(model, X_prev, fa): the two GPU providers (None = the library's); n_iter: refinement
    iterations of this step (default: configured); dropout = (p, seed, colonly) or None (v1 layout only)."""
    cfg = model._cfg
    H, L, flavour = cfg["hidden"], cfg["n_layers"], int(cfg.get("flavour", 0))
    n_iter = cfg["n_iter"] if n_iter is None else int(n_iter)
    if fa["X"].dim() != 3 or fa["X"].shape[1] != 1:
        raise ValueError("fabind_b200.train: X must be [N, 1, 3] (n_channel == 1, the published configuration)")
    if dropout is not None and dropout[0] > 0 and flavour == 1:
        raise NotImplementedError("fabind_b200.train: the FABind+ reverse pass carries no dropout masks")
    dev = fa["H"].device
    sd = state_dict if state_dict is not None else {k: v.detach() for k, v in model.state_dict().items()}
    X_prev = prev_coords(model, fa) if prev_coords is not None else _gpu_prev_coords(model, fa, n_iter, dropout)
    # the weight arena of this step does not depend on the coordinates: packed and converted BEFORE the edge lists are read back, so
    # that this host work (~1.3 ms) runs while the device is still busy with the earlier iterations (the edge lists have
    # data-dependent sizes: building them is the step's one synchronisation)
    packer = _packer(model, sd, H, L, flavour, dev) if (USE_FAST_PACKER and (dev.type == "cuda" or flavour == 0)) else None
    arena = packer.pack() if packer is not None else pack_state_dict(sd, H, L, flavour, device=dev)
    weights = slot_tensors(arena, H, L, flavour)
    lay = build_layout(fa["batch_id"], fa["segment_id"], fa["is_global"], fa["mask"], "cpu")
    ctx, inter = (edge_lists or _gpu_edges)(model, X_prev, fa)
    geo, edges, perm, moves = internal_graph(lay, ctx, inter, fa["compound_edge_index"], fa["LAS_edge_index"], dev)
    consts = dict(cmax=cfg["coord_clamp"], lcl=cfg["las_clamp"], las_step=cfg["las_step"], n_pairs=lay.P_total,
                  xl=fa["batched_complex_coord_LAS"].reshape(-1, 3)[perm].to(torch.float32).contiguous())
    if dropout is not None and dropout[0] > 0:
        consts["drop"] = bw.Drop(dropout[0], dropout[1], dropout[2], n_iter - 1)    # masks of the LAST (differentiated) iteration
    Hin = fa["H"].detach()[perm].to(torch.float32).contiguous()
    x_state = X_prev[:, 0][perm].to(torch.float32).contiguous()
    pair = None
    if flavour == 1:
        X_int, H_int, pair, tape, top = bw.stack_forward_train_plus(weights, Hin, x_state, moves, geo, edges, consts, L)
    else:
        X_int, H_int, tape, top = bw.stack_forward_train_v1(weights, Hin, x_state, moves, geo, edges, consts, L)
    X_out = torch.empty_like(fa["X"])
    X_out[perm, 0] = X_int.to(X_out.dtype)
    H_out = torch.empty(Hin.shape, dtype=fa["H"].dtype, device=dev)
    H_out[perm] = H_int.to(H_out.dtype)
    state = dict(H=H, L=L, flavour=flavour, sd=sd, arena=arena, weights=weights, tape=tape, top=top, geo=geo, edges=edges, consts=consts,
                 perm=perm, moves=moves, Hin_shape=Hin.shape, packer=packer, model=model)
    return X_out, H_out, pair, state


def _backward_half(st, gX, gH, gP=None):
    """steps 5: output gradients (caller order) -> ({parameter name: gradient}, dL/dH_in)"""
    perm, moves, flavour = st["perm"], st["moves"], st["flavour"]
    dX_int = (gX[:, 0][perm].to(torch.float32) * moves[:, None]).contiguous()
    dH_int = gH[perm].to(torch.float32).contiguous()
    garena = torch.zeros_like(st["arena"])
    slot_of = {name: (r, c, off) for name, r, c, off in slots(st["H"], st["L"], flavour)}
    pre_of = lambda name: name[:name.index(".") + 1] if "." in name else ""
    span = {}                                    # prefix -> [first element, end) of its run of slots
    for name, (r, c, off) in slot_of.items():
        lo, hi = span.get(pre_of(name), (off, off + r * c))
        span[pre_of(name)] = (min(lo, off), max(hi, off + r * c))
    ov = _overlap_state(st["model"], garena, max(hi for _, hi in span.values()))
    if flavour == 1:
        gP = torch.zeros(st["consts"]["n_pairs"], st["H"], dtype=torch.float32, device=dX_int.device) if gP is None else gP
        grads, dHin = bw.stack_backward_plus(st["weights"], st["tape"], st["top"], st["geo"], st["edges"], st["consts"], dH_int, dX_int,
                                             gP.to(torch.float32).contiguous())
        done = set()
    else:
        done = set()

        def on_group(pre, g):
            # the group's slots are contiguous in the arena (fb_weight_slot_*: one prefix = one run of slots): fill its slice now and,
            # with the overlapped all-reduce switched on, hand it to the collective while the reverse pass of the earlier layers runs
            for k, v in g.items():
                name = pre + k
                if name not in slot_of:
                    continue
                r, c, off = slot_of[name]
                garena[off:off + r * c] = v.reshape(-1)
                done.add(name)
            if ov is not None and pre in span:
                ov.reduce(pre, *span[pre])
        grads, dHin = bw.stack_backward_v1(st["weights"], st["tape"], st["top"], st["geo"], st["edges"], st["consts"], dH_int, dX_int,
                                           on_group=on_group)
    for name, (r, c, off) in slot_of.items():
        if name in grads and name not in done:
            garena[off:off + r * c] = grads[name].reshape(-1)
    if ov is not None:
        ov.finish(flavour)
    if st.get("packer") is not None:
        flat, pgrads = st["packer"].unpack(garena)
        # one flat fp32 buffer holding every parameter gradient in state_dict order: shard.allreduce_gradients reduces it in place
        # (no gather / scatter copies) when the parameters' .grad still alias it
        try:
            object.__setattr__(st["model"], "_fb_flat_grad", (flat, st["packer"].layout))
        except Exception:
            pass
    else:
        pgrads = arena_grads_to_state_dict(st["sd"], garena, st["H"], st["L"], flavour, device=garena.device)
    if ov is not None:
        # every parameter gradient of this step is already averaged over the ranks: shard.allreduce_gradients skips these parameters
        named = dict(st["model"].named_parameters())
        object.__setattr__(st["model"], "_fb_reduced", {id(named[k]) for k in pgrads if k in named})
    gH_in = torch.empty(st["Hin_shape"], dtype=torch.float32, device=dX_int.device)
    gH_in[perm] = dHin
    return pgrads, gH_in


class _Overlap:
    """Gradient all-reduce OVERLAPPED with the reverse pass (the role DDP's bucket hooks play for the reference, FABind/fabind/
    main_fabind.py:198-200).  The reverse pass finishes the weight gradients layer by layer, last layer first; each finished group is
    a contiguous slice of the ARENA gradient and is summed over the ranks on a side stream while the earlier layers are still being
    differentiated.  The arena -> state_dict chain rule (packer.unpack) is linear in the arena gradient and uses only the weights,
    which are identical on all ranks, so reducing before it gives the gradients of reducing after it (up to fp32 rounding order).
    Uncovered slots (none for the v1 layout) and the FABind+ layout are reduced in one tail collective."""

    def __init__(self, garena, group, average, end):
        import torch.distributed as dist
        self.dist, self.garena, self.group, self.average, self.end = dist, garena, group, average, end
        self.world = dist.get_world_size(group)
        self.handles, self.covered = [], []
        self.cuda = garena.device.type == "cuda"
        if self.cuda:
            self.main = torch.cuda.current_stream(garena.device)
            self.side = _side_stream(garena.device)

    def _issue(self, lo, hi):
        part = self.garena[lo:hi]
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(self.main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                self.handles.append(self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self.handles.append(self.dist.all_reduce(part, op=self.dist.ReduceOp.SUM, group=self.group, async_op=True))

    def reduce(self, pre, lo, hi):
        self.covered.append((lo, hi))
        self._issue(lo, hi)

    def finish(self, flavour):
        # whatever the hooks did not cover (gaps between the groups, the whole arena for the FABind+ layout) in one tail collective
        cur, gaps = 0, []
        for lo, hi in sorted(self.covered):
            if lo > cur:
                gaps.append((cur, lo))
            cur = max(cur, hi)
        if cur < self.end:                  # (the derived slots behind `end` carry no gradient)
            gaps.append((cur, self.end))
        for lo, hi in gaps:
            self._issue(lo, hi)
        for h in self.handles:
            h.wait()                       # CUDA: the side stream waits for the collective
        if self.cuda:
            self.main.wait_stream(self.side)
        if self.average:
            self.garena[:self.end].div_(self.world)


_SIDE = {}


def _side_stream(device):
    s = _SIDE.get(device)
    if s is None:
        s = _SIDE[device] = torch.cuda.Stream(device=device)
    return s


def overlap_allreduce(model, group=None, average=True, enabled=True):
    """Switch the overlapped gradient all-reduce of the training step on (or off) for `model` (an EfficientMCAttModel of this package
    used through `loss.backward()` or `training_step`): the collective then runs INSIDE the reverse pass, and a later
    `shard.allreduce_gradients(params, model=model)` only reduces parameters that do not belong to the stack."""
    object.__setattr__(model, "_fb_overlap", dict(group=group, average=average) if enabled else None)


def _overlap_state(model, garena, end):
    cfg = getattr(model, "_fb_overlap", None) if model is not None else None
    if cfg is None:
        return None
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(cfg["group"]) == 1:
        return None
    return _Overlap(garena, cfg["group"], cfg["average"], end)


def training_step(model, fa, output_grads, prev_coords=None, edge_lists=None, state_dict=None, n_iter=None, dropout=None):
    """model: fabind_b200.EfficientMCAttModel (v1 layout) or fabind_b200.plus.EfficientMCAttModel (FABind+ layout, eval-mode masks).
    fa: the forward arguments (X, H, batch_id, segment_id, mask, is_global, compound_edge_index, LAS_edge_index, batched_complex_coord_LAS).
    output_grads: v1  (X_out, H_out) -> (dL/dX_out, dL/dH_out);  FABind+  (X_out, H_out, pair rows [P,H]) -> (dL/dX, dL/dH, dL/dpair rows)
    (caller node order; pair rows packed per complex as [Np', Nc'] blocks).
    Returns (X_out, H_out[, pair rows], {parameter name: gradient}, dL/dH_in).  prev_coords / edge_lists: the two GPU providers
    (replaceable by their specifications in CPU tests)."""
    X_out, H_out, pair, st = _forward_half(model, fa, prev_coords, edge_lists, state_dict, n_iter, dropout)
    if st["flavour"] == 1:
        gX, gH, gP = output_grads(X_out, H_out, pair)
        pgrads, gH_in = _backward_half(st, gX, gH, gP)
        return X_out, H_out, pair, pgrads, gH_in
    gX, gH = output_grads(X_out, H_out)
    pgrads, gH_in = _backward_half(st, gX, gH)
    return X_out, H_out, pgrads, gH_in


class _StackFunction(torch.autograd.Function):
    """The differentiated last iteration as ONE autograd node: the reference's `loss.backward()` (main_fabind.py:380-401) reaches the
    stack's parameters and the incoming node features through it; everything inside runs on the library's kernels."""

    @staticmethod
    def forward(ctx, model, fa, prev_coords, edge_lists, names, step, H_in, *params):
        X_out, H_out, pair, st = _forward_half(model, fa, prev_coords, edge_lists, {n: p.detach() for n, p in zip(names, params)},
                                               step["n_iter"], step["dropout"])
        ctx.st, ctx.names, ctx.h_dtype = st, names, H_in.dtype
        ctx.pmeta = [(p.device, p.dtype) for p in params]
        if pair is None:
            return X_out, H_out
        return X_out, H_out, pair

    @staticmethod
    def backward(ctx, gX, gH, gP=None):
        # (autograd materialises undefined output gradients as zeros: gX / gH / gP are always tensors here)
        pgrads, gH_in = _backward_half(ctx.st, gX, gH, gP)
        ctx.st = None                                                   # the tape is consumed
        out = [None, None, None, None, None, None, gH_in.to(ctx.h_dtype)]
        for n, (dev, dt), p_needs in zip(ctx.names, ctx.pmeta, ctx.needs_input_grad[7:]):
            g = pgrads.pop(n)                                           # no second reference: autograd may adopt the view as .grad
            out.append(g.to(dev, dt) if p_needs else None)
        return tuple(out)


def forward_with_grad(model, fa, prev_coords=None, edge_lists=None, n_iter=None, dropout=None):
    """`EfficientMCAttModel.forward` with autograd: returns (X, H[, pair rows]) attached to the graph, so an unchanged training loop
    (`loss.backward()`, optimizer over `model.parameters()`) trains the drop-in module.  Buffers / non-float entries of the state_dict
    are passed through untouched.  n_iter: refinement iterations of this step; default = the reference's rule (att_model.py:210-211:
    `random.randint(1, n_iter)` in train() mode when --random-n-iter is set, else the configured n_iter).  dropout = (p, seed,
    colonly) or None.  The module's configuration is not mutated (re-entrant across streams / threads)."""
    full = model._cfg["n_iter"]
    if n_iter is None:
        n_iter = full
        if model.training and getattr(model, "random_n_iter", False):
            import random
            n_iter = random.randint(1, full)
    named = [(n, p) for n, p in model.state_dict(keep_vars=True).items() if torch.is_floating_point(p)]
    names = [n for n, _ in named]
    step = dict(n_iter=int(n_iter), dropout=dropout)
    out = _StackFunction.apply(model, fa, prev_coords, edge_lists, names, step, fa["H"], *[p for _, p in named])
    X = fa["X"]
    with torch.no_grad():
        X.copy_(out[0].detach())                 # the reference updates the caller's X in place (att_model.py:236,245)
    return out


training_step_v1 = training_step


def apply_gradients(model, pgrads, group=None, average=True):
    """parameter gradients -> `.grad` of the drop-in module's parameters, then ONE flat all-reduce over the ranks
    (`shard.allreduce_gradients`; unused parameters carry zeros, like DDP with find_unused_parameters)"""
    from .shard import allreduce_gradients
    params = dict(model.named_parameters())
    for k, p in params.items():
        g = pgrads.get(k)
        p.grad = None if g is None else g.to(p.device, p.dtype).reshape(p.shape)
    allreduce_gradients(list(params.values()), group=group, average=average, model=model)
