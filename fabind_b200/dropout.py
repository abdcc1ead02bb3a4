"""The dropout mask function of libfabind_b200 (csrc/common.cuh::fb_drop_hash), restated in numpy.

Masks are counter-based: keep(seed, site, row, col) is a pure function of the sampling seed (+ refinement iteration), the
nn.Dropout site of the reference (csrc/forward.cu::Site), the internal row id of the activation and the feature column.
This file is the specification the kernels are tested against (tests/emulate_packed.py, scripts/make_golden.py patch the
reference's nn.Dropout modules with `keep_mask(..., colonly=True)`); the product path never calls it.
"""
import numpy as np
import torch

SITES = dict(edge1=0, edge2=1, gcoord=2, node1=3, node2=4, patt=5, catt=6, ptr1=7, ptr2=8, ctr1=9, ctr2=10, pair1=11, pair2=12,
             agg=13, acoord=14, stack_in=15, stack_out=16)
M32 = np.uint64(0xFFFFFFFF)


def site_id(layer, name):
    """layer: -1 for the stack's own dropout, i for gcl_i / att_i, n_layers for out_layer"""
    return (layer + 1) * 32 + SITES[name]


def iter_seed(seed, it):
    return (int(seed) + 0x632BE5AB * int(it)) & 0xFFFFFFFF


def drop_hash(seed, site, row, col):
    """uint32 hash of broadcastable integer arrays (wrap-around arithmetic as in the CUDA code)"""
    row = np.asarray(row, dtype=np.uint64)
    col = np.asarray(col, dtype=np.uint64)
    x = np.uint64(seed & 0xFFFFFFFF) ^ ((np.uint64(site) * np.uint64(0x9E3779B1)) & M32)
    x = x ^ ((row * np.uint64(0x85EBCA77)) & M32)
    x = ((((x << np.uint64(13)) | (x >> np.uint64(19))) & M32) * np.uint64(0xC2B2AE3D)) & M32
    x = x ^ ((col * np.uint64(0x27D4EB2F)) & M32)
    x = x ^ (x >> np.uint64(15)); x = (x * np.uint64(0x2C1B3C6D)) & M32
    x = x ^ (x >> np.uint64(12)); x = (x * np.uint64(0x297A2D39)) & M32
    x = x ^ (x >> np.uint64(15))
    return x


def threshold(p):
    t = float(np.float32(p)) * 4294967296.0
    return 0xFFFFFFFF if t >= 4294967295.0 else int(t)


def keep_mask(seed, site, n_rows, n_cols, p, colonly=False, row0=0):
    """float32 [n_rows, n_cols] tensor: 1/(1-p) where the element is kept, 0 where it is dropped"""
    rows = np.zeros((n_rows, 1), dtype=np.uint64) if colonly else (np.arange(n_rows, dtype=np.uint64)[:, None] + np.uint64(row0))
    cols = np.arange(n_cols, dtype=np.uint64)[None, :]
    h = drop_hash(seed, site, rows, cols)
    h = np.broadcast_to(h, (n_rows, n_cols))
    scale = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return torch.from_numpy((h >= np.uint64(threshold(p))).astype(np.float32) * scale)
