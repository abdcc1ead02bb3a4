"""Deterministic, platform-independent weights for parity tests (TEST INFRASTRUCTURE).

The reference checkpoints are not available, so golden vectors are
produced with weights that can be regenerated anywhere from (key name, shape, seed) alone:
numpy PCG64 streams keyed by a CRC of the parameter name.  Matrices follow the fan-in-uniform
scale of torch.nn.Linear's default init; the final coordinate-head rows -- which the reference
initialises with xavier gain 0.001 (models/egnn.py:52,164; FABind+ models/egnn.py:41,135) and which would leave coordinates
frozen -- get an O(1) scale so clamps, the LAS step and the moving inter-edge set are exercised.
"""
import zlib

import numpy as np
import torch


def det_tensor(name, shape, seed=0):
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    shape = tuple(shape)
    if len(shape) == 2:
        fan_in = shape[1]
        bound = 1.0 / np.sqrt(fan_in)
        if name.endswith("coord_mlp.2.weight"):
            bound = 16.0 / np.sqrt(fan_in)
        if name.endswith("coord_mlp.linear2.weight"):   # FABind+: the head reads LayerNorm'd ReLU features (larger)
            bound = 3.0 / np.sqrt(fan_in)
    elif len(shape) == 1:
        bound = 0.05
    else:
        bound = 0.1
    t = torch.from_numpy(rng.uniform(-bound, bound, size=shape).astype(np.float32))
    if name.endswith("layernorm.weight"):
        t = 1.0 + t
    return t


def det_state_dict(shapes, seed=0):
    """shapes: mapping key -> shape (e.g. {k: v.shape for k, v in module.state_dict().items()})."""
    return {k: det_tensor(k, s, seed) for k, s in shapes.items()}
