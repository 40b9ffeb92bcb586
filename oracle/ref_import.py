"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference hot path.

The reference (/root/reference, read-only, build container only) is pure
Python/PyTorch.  Five third-party modules it imports at module scope are absent
here (hydra, omegaconf, torchsde, torchdiffeq, matplotlib); none of them does
arithmetic on the hot path, so they are stubbed in ``sys.modules`` before the
import.  Nothing here copies reference source.

``available()`` is False on the GPU box, where /root/reference does not exist;
callers must then rely on the committed fixtures under tests/golden/.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("BESO_REFERENCE_ROOT", "/root/reference")
_KD = "beso.agents.diffusion_agents.k_diffusion"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "beso", "agents", "diffusion_agents", "k_diffusion"))


def _instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    mod, name = target.rsplit(".", 1)
    cfg.update(kwargs)
    return getattr(importlib.import_module(mod), name)(*args, **cfg)


def _install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    hu = stub("hydra.utils", instantiate=_instantiate, call=_instantiate)
    stub("hydra", utils=hu)
    stub("omegaconf", DictConfig=dict, OmegaConf=object)
    stub("torchsde", BrownianTree=object)
    stub("torchdiffeq", odeint=None)
    plt = stub("matplotlib.pyplot")
    stub("matplotlib", pyplot=plt)


def load():
    """Returns a namespace with the reference modules score_wrappers, score_gpts,
    gc_sampling, classifier_free_sampler."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ns = types.SimpleNamespace()
    for name in ("score_wrappers", "score_gpts", "gc_sampling", "classifier_free_sampler", "utils"):
        setattr(ns, name, importlib.import_module(f"{_KD}.{name}"))
    return ns


def load_trajectory_loader():
    """The reference's beso/envs/dataloaders/trajectory_loader.py, loaded as a lone module by file path (its imports are
    torch / numpy / tqdm only; going through the package would pull in the simulators)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import importlib.util
    import itertools
    import torch._utils
    if not hasattr(torch._utils, "_accumulate"):   # private helper the module imports for its split function; removed
        torch._utils._accumulate = itertools.accumulate  # from current torch, never used on the slicing path
    path = os.path.join(REF_ROOT, "beso", "envs", "dataloaders", "trajectory_loader.py")
    spec = importlib.util.spec_from_file_location("_beso_ref_trajectory_loader", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_reference_model(ns, cfg, attn_pdrop=0.0, resid_pdrop=0.0, goal_drop=0.0):
    """Builds the reference GCDenoiser(DiffusionGPT) for an OracleCfg-like object."""
    inner = dict(
        _target_=f"{_KD}.score_gpts.DiffusionGPT",
        state_dim=cfg.obs_dim, device="cpu", goal_conditioned=cfg.goal_conditioned,
        action_dim=cfg.act_dim, embed_dim=cfg.d, embed_pdrob=0.0, attn_pdrop=attn_pdrop,
        resid_pdrop=resid_pdrop, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
        goal_seq_len=cfg.goal_len, obs_seq_len=cfg.window, sigma_vocab_size=0,
        time_embedding_fn=None, goal_drop=goal_drop, linear_output=cfg.linear_output)
    m = ns.score_wrappers.GCDenoiser(inner, sigma_data=cfg.sigma_data)
    m.eval()
    return m
