import ast
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Returns (ModelConfig, meta, dict of torch tensors) for tests/golden/<name>.npz."""
    from beso_b200.config import ModelConfig
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["meta"]))
    keys = ("obs_dim", "act_dim", "window", "goal_len", "d", "n_layers", "n_heads", "sigma_data",
            "linear_output", "goal_conditioned")
    cfg = ModelConfig(**{k: meta[k] for k in keys})
    arrays = {}
    for k in z.files:
        if k == "meta":
            continue
        a = z[k]
        arrays[k] = torch.from_numpy(a.copy()) if a.dtype.kind == "f" else a
    return cfg, meta, arrays


def golden_weights(cfg, meta):
    from beso_b200.synth import synthetic_state_dict
    sd = synthetic_state_dict(cfg, meta["weight_seed"])
    chk = float(sum(v.double().sum().item() for k, v in sd.items() if not k.endswith("attn.mask")))
    assert abs(chk - meta["weight_checksum"]) <= 1e-9 * max(1.0, abs(chk)), "synthetic weight generator drifted"
    return sd


def to_oracle_cfg(cfg):
    from oracle.beso_oracle import OracleCfg
    return OracleCfg(obs_dim=cfg.obs_dim, act_dim=cfg.act_dim, window=cfg.window, goal_len=cfg.goal_len,
                     d=cfg.d, n_layers=cfg.n_layers, n_heads=cfg.n_heads, sigma_data=cfg.sigma_data,
                     linear_output=cfg.linear_output, goal_conditioned=cfg.goal_conditioned)


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def load_checkpoint_golden(name):
    """tests/golden/ckpt_*.npz: trained weights of a shipped reference checkpoint (rounded to fp16 so that they can
    travel, widened back to fp32 here exactly as oracle/make_golden.py did before the reference loaded them) plus the
    reference's outputs.  Returns (ModelConfig, meta, arrays, state_dict without the attn.mask buffers)."""
    cfg, meta, arrays = load_golden(name)
    sd = {k[3:]: v.to(torch.float32) for k, v in arrays.items() if k.startswith("w::")}
    arrays = {k: v for k, v in arrays.items() if not k.startswith("w::")}
    return cfg, meta, arrays, sd


def with_masks(model, sd):
    """state_dict for a strict load: the checkpoint weights plus the model's own (deterministic) attn.mask buffers."""
    full = {k: v for k, v in model.state_dict().items() if k.endswith("attn.mask")}
    full.update(sd)
    return full
