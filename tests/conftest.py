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


def build_for_mode(cfg, device, mode="auto", **kw):
    """The tests' ``build_denoiser``.  Two extra mode names select the tile layout of the precise mode's tensor-core kernel
    through the library's test hook (``beso_debug_set_precise_layout``): ``precise128`` = 128-row tiles (three MMAs per
    product, embed_dim <= 256), ``precise64`` = stacked 64-row tiles; plain ``precise`` lets the launch choose (small test
    batches take the stacked layout, chip-filling batches the 128-row one).  The hook is global: it is set here for the
    model being built and reset after every test."""
    from beso_b200 import _lib
    from beso_b200.denoiser import build_denoiser
    if device is not None and str(device).startswith("cuda"):
        _lib.lib().beso_debug_set_precise_layout({"precise128": 2, "precise64": 1}.get(mode, 0))
    return build_denoiser(cfg, device, mode="precise" if mode in ("precise128", "precise64") else mode, **kw)


@pytest.fixture(autouse=True)
def _reset_precise_layout():
    yield
    import torch
    if torch.cuda.is_available():
        from beso_b200 import _lib
        _lib.lib().beso_debug_set_precise_layout(0)

