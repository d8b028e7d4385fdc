"""CPU restatement of the reference's MLP baseline (models/mlp.py:7-39) -- TEST INFRASTRUCTURE.

Only tests/ may import this module (see oracle/__init__.py).  numpy float64 throughout; the parameters come in
as the reference's ``state_dict`` (keys ``layers.<i>.weight`` ...).  Pinned against the unmodified reference by
tests/golden/mlp.npz (oracle/make_golden_mlp.py).

The dropout mask is NOT torch's (a different generator cannot be matched bit for bit): it is the counter-based
mask of the CUDA path -- keep element e iff u(seed, e) >= p with u = top 24 bits of
splitmix64(seed + e * 0x9E3779B97F4A7C15) / 2^24 (include/gsr_sm100.h, gsr_bn_act_fwd_f32) -- restated here with
numpy uint64 arithmetic so that the mask itself is checked bit exactly and everything downstream of it in float.
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def dropout_uniform(seed: int, count: int) -> np.ndarray:
    """u(seed, e) for e = 0..count-1, float32 in [0, 1)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.arange(count, dtype=np.uint64) * GOLDEN
        z ^= z >> np.uint64(30); z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27); z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def keep_mask(seed: int, shape, p: float) -> np.ndarray:
    if p <= 0:
        return np.ones(shape, dtype=bool)
    return (dropout_uniform(seed, int(np.prod(shape))) >= np.float32(p)).reshape(shape)


def layer_plan(state):
    """[(linear index, batchnorm index or None)] in forward order, from the state_dict keys (models/mlp.py:10-26:
    Flatten, then Linear, BatchNorm1d, Dropout, ReLU blocks, then the last Linear)."""
    lin = sorted(int(k.split(".")[1]) for k in state if k.endswith(".weight") and state[k].ndim == 2)
    bn = {int(k.split(".")[1]) for k in state if k.endswith(".running_mean")}
    return [(i, i + 1 if i + 1 in bn else None) for i in lin]


def forward(state, x, training=False, p=0.1, seed=0, eps=1e-5, momentum=0.1, return_cache=False):
    """models/mlp.py:33-39.  Returns (y [B, h, h], new running statistics {key: array}); with ``return_cache`` also
    the per-block intermediates for ``backward``."""
    x = np.asarray(x, dtype=np.float64).reshape(len(x), -1)                       # nn.Flatten(start_dim=1)
    new_stats, cache = {}, []
    for li, bi in layer_plan(state):
        W, b = np.asarray(state[f"layers.{li}.weight"], np.float64), np.asarray(state[f"layers.{li}.bias"], np.float64)
        a = x @ W.T + b                                                           # nn.Linear
        if bi is None:
            cache.append(dict(x=x, li=li))
            x = a
            continue
        g, be = np.asarray(state[f"layers.{bi}.weight"], np.float64), np.asarray(state[f"layers.{bi}.bias"], np.float64)
        if training:                                                              # nn.BatchNorm1d, batch statistics
            B = a.shape[0]
            mean, var = a.mean(0), a.var(0)
            rm = np.asarray(state[f"layers.{bi}.running_mean"], np.float64)
            rv = np.asarray(state[f"layers.{bi}.running_var"], np.float64)
            new_stats[f"layers.{bi}.running_mean"] = (1 - momentum) * rm + momentum * mean
            new_stats[f"layers.{bi}.running_var"] = (1 - momentum) * rv + momentum * var * B / max(B - 1, 1)
        else:
            mean = np.asarray(state[f"layers.{bi}.running_mean"], np.float64)
            var = np.asarray(state[f"layers.{bi}.running_var"], np.float64)
        rstd = 1.0 / np.sqrt(var + eps)
        xhat = (a - mean) * rstd
        z = xhat * g + be
        keep = keep_mask(seed, z.shape, p if training else 0.0)                   # nn.Dropout(p)
        scale = 1.0 / (1.0 - p) if (training and p > 0) else 1.0
        d = np.where(keep, z * scale, 0.0)
        y = np.maximum(d, 0.0)                                                    # nn.ReLU
        cache.append(dict(x=x, li=li, bi=bi, xhat=xhat, rstd=rstd, keep=keep, scale=scale, d=d, training=training))
        x = y
    h = int(np.sqrt(x.shape[1]))
    out = x.reshape(len(x), h, h)                                                 # nn.Unflatten
    return (out, new_stats, cache) if return_cache else (out, new_stats)


def backward(state, cache, grad_out):
    """Gradients of sum(grad_out * y) w.r.t. every parameter (keys as the state_dict) and the input."""
    g = np.asarray(grad_out, np.float64).reshape(len(grad_out), -1)
    grads = {}
    for c in reversed(cache):
        li = c["li"]
        W = np.asarray(state[f"layers.{li}.weight"], np.float64)
        if "bi" in c:
            bi = c["bi"]
            gamma = np.asarray(state[f"layers.{bi}.weight"], np.float64)
            g = np.where((c["d"] > 0) & c["keep"], g * c["scale"], 0.0)
            grads[f"layers.{bi}.bias"] = g.sum(0)
            grads[f"layers.{bi}.weight"] = (g * c["xhat"]).sum(0)
            if c["training"]:
                B = g.shape[0]
                g = g - grads[f"layers.{bi}.bias"] / B - c["xhat"] * grads[f"layers.{bi}.weight"] / B
            g = gamma * c["rstd"] * g
        grads[f"layers.{li}.bias"] = g.sum(0)
        grads[f"layers.{li}.weight"] = g.T @ c["x"]
        g = g @ W
    return grads, g
