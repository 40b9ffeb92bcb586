"""Deterministic synthetic weights and inputs (no network, no checkpoints).

``synthetic_state_dict`` draws every tensor from a numpy PCG64 stream keyed by
(seed, parameter index), so the build container (where the golden fixtures are
made with the real reference) and the GPU box regenerate identical weights
without shipping them.  The distribution follows the reference init
(score_gpts.py:202-211: Linear/pos_emb ~ N(0, 0.02)) but, unlike it, biases and
LayerNorm affine terms are non-trivial so every term of the forward is exercised.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .config import ModelConfig


def synthetic_state_dict(cfg: ModelConfig, seed: int = 0, weight_std: float = 0.02,
                         with_buffers: bool = True) -> "OrderedDict[str, torch.Tensor]":
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, (name, shape) in enumerate(cfg.param_shapes()):
        rng = np.random.Generator(np.random.PCG64([seed, i]))
        x = rng.standard_normal(shape, dtype=np.float32)
        if name.endswith("ln1.weight") or name.endswith("ln2.weight") or name.endswith("ln_f.weight"):
            x = 1.0 + 0.1 * x
        elif name.endswith(".bias"):
            x = 0.02 * x
        elif name.endswith("sigma_emb.weight"):
            x = 0.2 * x
        else:
            x = weight_std * x
        sd[name] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        if with_buffers and name.endswith("ln2.bias"):
            # persistent causal-mask buffer, present in the reference state_dict right
            # after ln2.bias (score_gpts.py:42-47)
            bs = cfg.block_size
            sd[name.replace("ln2.bias", "attn.mask")] = torch.tril(torch.ones(bs, bs)).view(1, 1, bs, bs)
    return sd


def synthetic_inputs(cfg: ModelConfig, batch: int, seed: int = 0, t: int | None = None,
                     sigma_min: float = 0.005, sigma_max: float = 1.0):
    """state (B,t,obs) ~ N(0,1), goal (B,G,obs) ~ N(0,1), clean action ~ U(-1,1),
    noise ~ N(0,1), sigma ~ log-uniform[sigma_min, sigma_max]; action = clean + sigma*noise."""
    t = cfg.window if t is None else t
    rng = np.random.Generator(np.random.PCG64([seed, 10_000]))
    f32 = np.float32
    state = rng.standard_normal((batch, t, cfg.obs_dim), dtype=f32)
    goal = rng.standard_normal((batch, cfg.goal_len, cfg.obs_dim), dtype=f32)
    clean = rng.uniform(-1.0, 1.0, (batch, t, cfg.act_dim)).astype(f32)
    noise = rng.standard_normal((batch, t, cfg.act_dim), dtype=f32)
    u = rng.uniform(0.0, 1.0, (batch,))
    sigma = np.exp(np.log(sigma_min) + u * (np.log(sigma_max) - np.log(sigma_min))).astype(f32)
    action = (clean + sigma[:, None, None] * noise).astype(f32)
    tt = torch.from_numpy
    return dict(state=tt(state), goal=tt(goal), clean=tt(clean), noise=tt(noise),
                sigma=tt(sigma), action=tt(action))
