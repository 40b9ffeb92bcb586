"""Drop-in ``GCDenoiser`` / ``DiffusionGPT`` whose forward is ONE sm_100a kernel launch.

Mirrors the interface the reference agent shell uses:

* ``GCDenoiser(inner_model, sigma_data)``        k_diffusion/score_wrappers.py:18-99
* ``DiffusionGPT(state_dim, device, ...)``        k_diffusion/score_gpts.py:118-374

The modules below hold ``nn.Parameter``s with exactly the reference's names, shapes,
``parameters()`` order and ``state_dict`` schema (including the persistent
``blocks.N.attn.mask`` buffers), so checkpoints, optimizers and the EMA helper work unchanged.
They contain no PyTorch arithmetic: ``forward`` hands raw device pointers to
``libbeso_b200.so`` (include/beso_b200.h).  Without the library, or on a non-CUDA tensor,
calls raise -- there is no eager fallback.
"""
from __future__ import annotations

import ctypes as C
import importlib
import weakref
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .config import ModelConfig


def _instantiate(spec, **overrides):
    """hydra.utils.instantiate when Hydra is installed, else the same thing for a plain
    mapping with a ``_target_`` (score_wrappers.py:28)."""
    try:  # pragma: no cover - hydra is not in the build image
        import hydra
        return hydra.utils.instantiate(spec, **overrides)
    except ImportError:
        spec = dict(spec)
        target = spec.pop("_target_")
        spec.pop("_recursive_", None)
        spec.update(overrides)
        mod, name = target.rsplit(".", 1)
        return getattr(importlib.import_module(mod), name)(**spec)


class _Attn(nn.Module):
    """Parameter container for CausalSelfAttention (score_gpts.py:15-48)."""

    def __init__(self, d: int, n_heads: int, attn_pdrop: float, resid_pdrop: float, block_size: int):
        super().__init__()
        if d % n_heads:
            raise AssertionError("embed_dim must be divisible by n_heads")
        self.key = nn.Linear(d, d)
        self.query = nn.Linear(d, d)
        self.value = nn.Linear(d, d)
        self.attn_drop = nn.Dropout(attn_pdrop)
        self.resid_drop = nn.Dropout(resid_pdrop)
        self.proj = nn.Linear(d, d)
        self.register_buffer("mask", torch.ones(block_size, block_size).tril_().view(1, 1, block_size, block_size))
        self.n_head = n_heads


class _Block(nn.Module):
    """Parameter container for Block (score_gpts.py:83-110)."""

    def __init__(self, d: int, n_heads: int, attn_pdrop: float, resid_pdrop: float, block_size: int):
        super().__init__()
        self.ln1 = nn.LayerNorm(d)
        self.ln2 = nn.LayerNorm(d)
        self.attn = _Attn(d, n_heads, attn_pdrop, resid_pdrop, block_size)
        self.mlp = nn.Sequential(nn.Linear(d, 4 * d), nn.GELU(), nn.Linear(4 * d, d), nn.Dropout(resid_pdrop))


class DiffusionGPT(nn.Module):
    """Score-GPT over ``[sigma, g_1..g_G, s_1, a_1, ..., s_t, a_t]`` (score_gpts.py:118-374)."""

    def __init__(self, state_dim: int, device: str, goal_conditioned: bool, action_dim: int, embed_dim: int,
                 embed_pdrob: float, attn_pdrop: float, resid_pdrop: float, n_layers: int, n_heads: int,
                 goal_seq_len: int, obs_seq_len: int, sigma_vocab_size: int = 0, time_embedding_fn=None,
                 goal_drop: float = 0, linear_output: bool = False):
        super().__init__()
        self.device = device
        self.goal_conditioned = goal_conditioned
        if not goal_conditioned:
            goal_seq_len = 0
        self.block_size = goal_seq_len + 2 * obs_seq_len + 1
        seq_size = goal_seq_len + obs_seq_len + 1
        self.tok_emb = nn.Linear(state_dim, embed_dim)
        self.pos_emb = nn.Parameter(torch.zeros(1, seq_size, embed_dim))
        self.drop = nn.Dropout(embed_pdrob)
        self.cond_mask_prob = goal_drop
        self.action_dim, self.obs_dim, self.embed_dim = action_dim, state_dim, embed_dim
        self.blocks = nn.Sequential(*[_Block(embed_dim, n_heads, attn_pdrop, resid_pdrop, self.block_size)
                                      for _ in range(n_layers)])
        self.ln_f = nn.LayerNorm(embed_dim)
        self.goal_seq_len, self.obs_seq_len = goal_seq_len, obs_seq_len
        self.sigma_emb = nn.Linear(1, embed_dim)
        self.action_emb = nn.Linear(action_dim, embed_dim)
        if linear_output:
            self.action_pred = nn.Linear(embed_dim, action_dim)
        else:
            self.action_pred = nn.Sequential(nn.Linear(embed_dim, 100), nn.SiLU(), nn.Linear(100, action_dim))
        self._dropouts = (embed_pdrob, attn_pdrop, resid_pdrop)
        self.config = ModelConfig(obs_dim=state_dim, act_dim=action_dim, window=obs_seq_len,
                                  goal_len=goal_seq_len, d=embed_dim, n_layers=n_layers, n_heads=n_heads,
                                  sigma_data=1.0, linear_output=bool(linear_output),
                                  goal_conditioned=bool(goal_conditioned))
        self.reset_parameters()

    def reset_parameters(self):
        """Same distributions as score_gpts.py:202-211: Linear and pos_emb ~ N(0, 0.02), biases 0,
        LayerNorm (1, 0)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.pos_emb, mean=0.0, std=0.02)

    def get_block_size(self):
        return self.block_size

    def get_params(self):
        return self.parameters()

    def mask_cond(self, cond: torch.Tensor, force_mask: bool = False) -> torch.Tensor:
        """Element-wise Bernoulli goal masking for CFG training (score_gpts.py:360-371)."""
        if force_mask:
            return torch.zeros_like(cond)
        if self.training and self.cond_mask_prob > 0.0:
            mask = torch.bernoulli(torch.ones(cond.shape, device=cond.device) * self.cond_mask_prob)
            return cond * (1.0 - mask)
        return cond

    def forward(self, states, actions, goals, sigma, uncond: Optional[bool] = False,
                keep_last_actions: Optional[bool] = False):
        """DiffusionGPT.forward (score_gpts.py:272-358) without the Karras pre-conditioning."""
        ref = self.__dict__.get("_owner_ref")
        owner = ref() if ref is not None else None
        if owner is None:
            owner = GCDenoiser(self, sigma_data=1.0)
            self.__dict__["_standalone_owner"] = owner            # keep the plan alive
        return owner._run(states, actions, goals, sigma, uncond=bool(uncond), inner=True,
                          keep_last_actions=bool(keep_last_actions))


class GCDenoiser(nn.Module):
    """Karras et al. pre-conditioner around the score-GPT (score_wrappers.py:18-99).

    ``mode``: "precise" (fp32-equivalent arithmetic, rtol 1e-3 / atol 1e-5 against the fp32 reference: the
    split-operand tcgen05 kernel where the shape is supported, else the fp32 CUDA-core kernel), "fast" (fp16
    tcgen05 kernel, single pass: 3.7x the throughput, max |err| ~1e-3 on random-init and ~3e-3 on trained weights:
    an explicit opt-in), "simt" (always the fp32 CUDA-core kernel).  "auto" (the default) is "precise": a drop-in for
    the reference has to reproduce its results within the stated tolerance before it is fast.
    """

    def __init__(self, inner_model, sigma_data: float = 1.0, mode: str = "auto"):
        super().__init__()
        self.inner_model = inner_model if isinstance(inner_model, nn.Module) else _instantiate(inner_model)
        self.sigma_data = sigma_data
        self.mode = mode
        self.inner_model.__dict__["_owner_ref"] = weakref.ref(self)   # no module cycle, deepcopy-safe
        self.inner_model.__dict__.pop("_standalone_owner", None)
        self._plan = None
        self._plan_key = None
        self._packed = {}          # slot -> fingerprint
        self._slot = 0

    # ---- reference API -------------------------------------------------------------------
    def get_scalings(self, sigma):
        """score_wrappers.py:31-43"""
        c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
        c_out = sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in

    def get_params(self):
        return self.inner_model.parameters()

    def forward(self, state, action, goal, sigma, **kwargs):
        """score_wrappers.py:81-96; kwargs: uncond, keep_last_actions."""
        unknown = set(kwargs) - {"uncond", "keep_last_actions"}
        if unknown:
            raise TypeError(f"unexpected keyword arguments {sorted(unknown)}")
        return self._run(state, action, goal, sigma, uncond=bool(kwargs.get("uncond", False)), inner=False,
                         keep_last_actions=bool(kwargs.get("keep_last_actions", False)))

    def loss(self, state, action, goal, noise, sigma, **kwargs):
        """score_wrappers.py:45-79.  Returns a scalar tensor whose backward fills ``.grad`` of every
        parameter (hand-written backward kernels, see beso_b200/training.py)."""
        from .training import denoiser_loss
        return denoiser_loss(self, state, action, goal, noise, sigma, **kwargs)

    # ---- plan / weights ------------------------------------------------------------------
    @property
    def config(self) -> ModelConfig:
        c = self.inner_model.config
        return ModelConfig(**{**c.__dict__, "sigma_data": float(self.sigma_data)})

    def _device_index(self, t: torch.Tensor) -> int:
        if not t.is_cuda:
            raise _lib.BesoLibraryError(
                "beso_b200 runs on CUDA tensors only (there is no CPU/PyTorch fallback); got a "
                f"{t.device} tensor")
        return t.device.index if t.device.index is not None else torch.cuda.current_device()

    def _ensure_plan(self, dev: int):
        cfg = self.config
        key = (cfg, dev)
        if self._plan is None or self._plan_key != key:
            self.close()
            handle = C.c_void_p()
            desc = _lib.ModelDesc.from_config(cfg)
            _lib.check(_lib.lib().beso_plan_create(C.byref(desc), dev, C.byref(handle)), "beso_plan_create")
            self._plan, self._plan_key, self._packed = handle, key, {}
        return self._plan

    def _params(self):
        """The inner model's parameters in ``parameters()`` order WITHOUT walking the module tree on every call (the walk
        costs more host time than a whole batch-1 forward takes on the GPU): the ``(module._parameters, name)`` slots are
        cached, so a re-assigned Parameter object is still seen; the cache is dropped by ``_apply`` (``.to`` / ``.cuda``
        / ``.float``), ``load_state_dict`` and when ``inner_model`` is replaced."""
        slots = self.__dict__.get("_pslots")
        if slots is None:
            slots, seen = [], set()
            for mod in self.inner_model.modules():
                for name, prm in mod._parameters.items():
                    if prm is not None and id(prm) not in seen:
                        seen.add(id(prm))
                        slots.append((mod._parameters, name))
            if [id(d[n]) for d, n in slots] != [id(q) for q in self.inner_model.parameters()]:      # pragma: no cover
                raise _lib.BesoLibraryError("internal: cached parameter slots are not in parameters() order")
            self.__dict__["_pslots"] = slots
        return [d[n] for d, n in slots]

    def _drop_param_cache(self):
        self.__dict__.pop("_pslots", None)

    def _apply(self, fn, *args, **kwargs):
        self._drop_param_cache()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._drop_param_cache()
        return super().load_state_dict(*args, **kwargs)

    def __setattr__(self, name, value):
        if name == "inner_model":
            self.__dict__.pop("_pslots", None)
        super().__setattr__(name, value)

    def _fingerprint(self, params=None):
        return tuple((p.data_ptr(), p._version) for p in (self._params() if params is None else params))

    def refresh_weights(self, slot: Optional[int] = None, force: bool = True):
        """(Re)pack the current parameter values into weight slot ``slot`` (0 = raw, 1 = EMA) and
        select it.  Called automatically when a parameter's storage or version counter changes;
        call it by hand after writing through ``param.data`` (e.g. the reference EMA helper's
        ``copy_to`` / ``restore``), which PyTorch's version counters do not see."""
        slot = self._slot if slot is None else slot
        params = self._params()
        if not params or not params[0].is_cuda:
            raise _lib.BesoLibraryError("model parameters must live on a CUDA device (call .to('cuda'))")
        dev = self._device_index(params[0])
        plan = self._ensure_plan(dev)
        # Slot 1 holds weights handed over explicitly (pack_tensors: the EMA copy).  It is never re-packed behind
        # the caller's back from the live parameters: a stale fingerprint there would silently replace the EMA
        # weights by the raw ones.  Only an explicit refresh_weights(slot=1, force=True) packs the live values into it.
        external = slot in self._packed and self._packed[slot][0] == "external"
        fp = self._fingerprint(params)
        if force or (not external and self._packed.get(slot) != fp):
            self._pack(plan, dev, slot, params)
            self._packed[slot] = fp
        _lib.check(_lib.lib().beso_plan_select_weights(plan, slot), "beso_plan_select_weights")
        self._slot = slot
        return plan

    def _pack(self, plan, dev, slot, tensors):
        for p in tensors:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.is_cuda:
                raise _lib.BesoLibraryError("weights must be contiguous fp32 CUDA tensors")
        ptrs = (C.c_void_p * len(tensors))(*[p.data_ptr() for p in tensors])
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().beso_plan_pack_weights(plan, slot, ptrs, len(tensors), C.c_void_p(stream)),
                   "beso_plan_pack_weights")

    def pack_tensors(self, slot: int, tensors, tag=None):
        """Pack ``tensors`` (parameters() order and shapes; e.g. the EMA shadow copy) into weight slot ``slot`` straight
        from where they live -- the live parameters are not touched -- and mark the slot as externally owned: it is
        re-packed only by another ``pack_tensors`` call, never automatically.  ``tag`` is remembered for the caller's
        staleness bookkeeping (``packed_tag``)."""
        params = list(self.inner_model.parameters())
        tensors = [t.detach() for t in tensors]
        if len(tensors) != len(params) or any(t.numel() != p.numel() for t, p in zip(tensors, params)):
            raise ValueError("pack_tensors: tensors must match parameters() in number and size")
        dev = self._device_index(params[0])
        plan = self._ensure_plan(dev)
        self._pack(plan, dev, slot, tensors)
        self._packed[slot] = ("external", tag)

    def packed_tag(self, slot: int):
        e = self._packed.get(slot)
        return e[1] if e is not None and e[0] == "external" else None

    def select_weights(self, slot: int):
        """Switch between resident packed weight sets without re-packing."""
        _lib.check(_lib.lib().beso_plan_select_weights(self._plan, slot), "beso_plan_select_weights")
        self._slot = slot

    def fast_supported(self) -> bool:
        """Whether the tensor-core kernel takes this model's shape -- asked of the library (the one place that
        knows the kernel's limits), never re-derived here."""
        if self._plan is None:
            p0 = next(self.inner_model.parameters())
            if not p0.is_cuda:
                return False
            self._ensure_plan(self._device_index(p0))
        return _lib.lib().beso_plan_rows_per_cta(self._plan, _lib.MODE_FAST, self.config.window) > 0

    def resolved_mode(self) -> int:
        if self.mode == "auto":
            return _lib.MODE_PRECISE
        return _lib.MODE_IDS[self.mode]

    def close(self):
        if self._plan is not None:
            _lib.lib().beso_plan_destroy(self._plan)
            self._plan, self._plan_key, self._packed = None, None, {}

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __deepcopy__(self, memo):
        """The agent shell deep-copies models (base_workspace_manager.py:296,373); the plan handle
        is per-instance and rebuilt lazily."""
        import copy
        inner = copy.deepcopy(self.inner_model, memo)
        new = GCDenoiser(inner, sigma_data=self.sigma_data, mode=self.mode)
        new.train(self.training)
        for k, v in self.__dict__.items():
            if k not in new.__dict__ and not k.startswith("_"):
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    # ---- launch --------------------------------------------------------------------------
    @staticmethod
    def _prep(x: torch.Tensor) -> torch.Tensor:
        if x.dtype != torch.float32:
            x = x.float()
        return x if x.is_contiguous() else x.contiguous()

    def _check_shapes(self, state, action, goal):
        cfg = self.config
        if state.dim() != 3 or action.dim() != 3:
            raise ValueError("state and action must be (B, t, dim)")
        B, t, _ = state.shape
        assert t <= self.inner_model.block_size, "Cannot forward, model block size is exhausted."
        if t > cfg.window:
            raise ValueError(f"t={t} exceeds obs_seq_len={cfg.window}")
        if state.shape[2] != cfg.obs_dim or tuple(action.shape) != (B, t, cfg.act_dim):
            raise ValueError("state/action shapes do not match the model")
        if cfg.G and tuple(goal.shape) != (B, cfg.G, cfg.obs_dim):
            raise ValueError(f"goal must be (B, {cfg.G}, {cfg.obs_dim}), got {tuple(goal.shape)}")
        return B, t

    def _run(self, state, action, goal, sigma, uncond=False, inner=False, keep_last_actions=False,
             cfg_lambda: Optional[float] = None):
        dev = self._device_index(action)
        if self.inner_model.training and torch.is_grad_enabled() and any(
                p.requires_grad for p in self.inner_model.parameters()):
            raise _lib.BesoLibraryError(
                "GCDenoiser.forward under autograd is not supported; use .loss() for training or "
                "torch.no_grad() for sampling")
        if self.inner_model.training and (self.inner_model.cond_mask_prob > 0 and not uncond):
            goal = self.inner_model.mask_cond(goal)              # score_gpts.py:298-299
        state, action, goal, sigma = map(self._prep, (state, action, goal, sigma))
        io = self.__dict__.get("io_scaling")
        if io is not None:                      # step-wise (Python loop) sampling under rollout scaling: raw inputs
            state, goal = io.scale_inputs(state, goal)
        B, t = self._check_shapes(state, action, goal)
        if sigma.dim() == 0:
            sigma = sigma.expand(B).contiguous()
        plan = self.refresh_weights(force=False)
        flags = (_lib.FLAG_UNCOND if uncond else 0) | (_lib.FLAG_INNER if (inner or keep_last_actions) else 0)
        lam = 0.0
        if cfg_lambda is not None:
            flags |= _lib.FLAG_CFG
            lam = float(cfg_lambda)
        x_in = action
        if keep_last_actions and not inner:
            x_in = action * self.get_scalings(sigma)[2].view(-1, 1, 1)
        out = torch.empty_like(action)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().beso_denoise_fwd(plan, self.resolved_mode(), state.data_ptr(), x_in.data_ptr(),
                                              goal.data_ptr(), sigma.data_ptr(), out.data_ptr(), B, t, flags,
                                              lam, C.c_void_p(stream)), "beso_denoise_fwd")
        if keep_last_actions:                                     # score_gpts.py:355-356 (B == 1 only)
            out = torch.cat([x_in[:, :-1, :], out[:, -1, :].reshape(1, 1, -1)], dim=1)
            if not inner:
                c_skip, c_out, _ = self.get_scalings(sigma)
                out = out * c_out.view(-1, 1, 1) + action * c_skip.view(-1, 1, 1)
        return out

    def sample(self, sampler: str, sigmas: torch.Tensor, state, x_t, goal, cfg_lambda: Optional[float] = None,
               uncond: bool = False, coef: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The whole DDIM / Euler / Heun / Euler-ancestral loop as one persistent kernel launch
        (beso_sample_loop_scaled).  ``noise``: (n_steps, B, t, act) standard-normal draws for the ancestral
        sampler.  With ``self.io_scaling`` set (``RolloutScaling``; the agent does for predict / evaluate) ``state`` and
        ``goal`` are RAW observations: the kernel scales them on its first read, clips the final actions and writes
        their inverse-scaled copy to ``io_scaling.unscaled``."""
        dev = self._device_index(x_t)
        state, x_t, goal = map(self._prep, (state, x_t, goal))
        B, t = self._check_shapes(state, x_t, goal)
        plan = self.refresh_weights(force=False)
        sig = [float(v) for v in sigmas.detach().cpu().float().tolist()]
        sig_arr = _lib.float_array(sig)
        coef_arr = _lib.float_array(coef.detach().cpu().float().reshape(-1).tolist()) if coef is not None else None
        flags = _lib.FLAG_UNCOND if uncond else 0
        lam = 0.0
        if cfg_lambda is not None:
            flags |= _lib.FLAG_CFG
            lam = float(cfg_lambda)
        x = x_t.clone()                                           # samplers never write the caller's x_t
        noise_ptr = None
        if noise is not None:
            noise = self._prep(noise)
            if tuple(noise.shape) != (len(sig) - 1,) + tuple(x.shape) or noise.device != x.device:
                raise ValueError(f"noise must be {(len(sig) - 1,) + tuple(x.shape)} on {x.device}")
            noise_ptr = noise.data_ptr()
        io = self.__dict__.get("io_scaling")
        io_struct = None
        if io is not None:
            io.unscaled = torch.empty_like(x)
            io_struct = io.struct(io.unscaled)
            io.consumed = True
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.lib().beso_sample_loop_scaled(plan, self.resolved_mode(), _lib.SAMPLER_IDS[sampler], sig_arr,
                                                     len(sig), coef_arr, state.data_ptr(), goal.data_ptr(),
                                                     x.data_ptr(), noise_ptr,
                                                     C.byref(io_struct) if io_struct is not None else None, B, t, flags,
                                                     lam, C.c_void_p(stream)),
                   "beso_sample_loop_scaled")
        return x


class RolloutScaling:
    """The scaler calls around ``sample_loop`` in ``BesoAgent.predict`` / ``evaluate`` (beso_agent.py:322-329,373-387,
    base_agent.py:111-142) as tables for the sampling kernel (``beso_io_scaling``): ``scale_input`` of states and goals,
    the zeroed block-push goal dimensions, ``clip_action`` and ``inverse_scale_output``.

    ``in_table`` / ``out_table``: (4, dim) fp32 device tensors with rows (sub, div, mul, add); ``goal_keep``: (obs,) fp32;
    ``clip``: (2, act) float64.  After a fused loop ``consumed`` is True and ``unscaled`` holds the clipped,
    inverse-scaled actions; a step-wise (Python) loop leaves ``consumed`` False and the caller applies clip / inverse
    scaling itself (``scale_inputs`` is applied by the model's forward in that case)."""

    def __init__(self, in_table=None, goal_keep=None, clip=None, out_table=None):
        self.in_table, self.goal_keep, self.clip, self.out_table = in_table, goal_keep, clip, out_table
        self.unscaled = None
        self.consumed = False

    def struct(self, unscaled):
        s = _lib.IoScaling()
        s.in_table = self.in_table.data_ptr() if self.in_table is not None else None
        s.goal_keep = self.goal_keep.data_ptr() if self.goal_keep is not None else None
        s.out_clip = self.clip.data_ptr() if self.clip is not None else None
        s.out_table = self.out_table.data_ptr() if self.out_table is not None else None
        s.unscaled_out = unscaled.data_ptr()
        return s

    @staticmethod
    def _apply(x, tab):
        return ((x - tab[0]) / tab[1]) * tab[2] + tab[3]

    def scale_inputs(self, state, goal):
        if self.in_table is not None:
            state, goal = self._apply(state, self.in_table), self._apply(goal, self.in_table)
        if self.goal_keep is not None:
            goal = goal * self.goal_keep
        return state, goal

    def finish(self, x):
        """clip + inverse scaling in torch, for the step-wise path."""
        if self.clip is not None:
            x = torch.clamp(x, self.clip[0], self.clip[1]).to(torch.float32)
        y = self._apply(x, self.out_table) if self.out_table is not None else x
        return x, y


def build_denoiser(cfg: ModelConfig, device="cuda", mode: str = "auto", state_dict=None,
                   attn_pdrop: float = 0.0, resid_pdrop: float = 0.0, goal_drop: float = 0.0) -> GCDenoiser:
    """Convenience constructor taking a ``ModelConfig`` instead of the Hydra argument list."""
    inner = DiffusionGPT(state_dim=cfg.obs_dim, device=str(device), goal_conditioned=cfg.goal_conditioned,
                         action_dim=cfg.act_dim, embed_dim=cfg.d, embed_pdrob=0.0, attn_pdrop=attn_pdrop,
                         resid_pdrop=resid_pdrop, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                         goal_seq_len=cfg.goal_len, obs_seq_len=cfg.window, sigma_vocab_size=0,
                         time_embedding_fn=None, goal_drop=goal_drop, linear_output=cfg.linear_output)
    model = GCDenoiser(inner, sigma_data=cfg.sigma_data, mode=mode)
    if state_dict is not None:
        model.load_state_dict(state_dict, strict=True)
    model.eval()
    return model.to(device)
