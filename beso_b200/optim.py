"""Fused optimiser step for the training loop.

``FusedAdamW`` is a ``torch.optim.Optimizer`` (so ``torch.optim.lr_scheduler.StepLR`` and the reference's
``BesoAgent.train_step`` -- beso_agent.py:238-247 -- work unchanged) whose ``step()`` is ONE launch of
``beso_opt_step`` over all parameter tensors: AdamW exactly as ``torch.optim.AdamW`` computes it
(configs/agents/beso_kitchen.yaml:9-12) and, when an ``ExponentialMovingAverage`` from this module is
attached, the EMA update of beso/networks/ema_helper/ema.py:36-53 in the same pass.

There is no PyTorch fallback: parameters must live on a CUDA device and the library must be built.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import torch

from . import _lib


def ema_decay_at(decay: float, num_updates: Optional[int]) -> float:
    """Decay used by the n-th update (n = num_updates AFTER the increment), ema.py:46-50."""
    if num_updates is None:
        return decay
    return min(decay, (1 + num_updates) / (10 + num_updates))


class ExponentialMovingAverage:
    """Mirror of beso/networks/ema_helper/ema.py with the shadow parameters in ONE flat buffer.

    Attached to a ``FusedAdamW`` (``opt.attach_ema(ema)``) the update is applied inside the optimiser's
    kernel and ``update()`` only acknowledges it; stand-alone use is refused (no PyTorch fallback)."""

    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float, device: str = "cuda", use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        params = [p for p in parameters if p.requires_grad]
        self.decay = decay
        self._device = device
        self.num_updates = 0 if use_num_updates else None
        self.flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
        self.shadow_params, off = [], 0
        for p in params:                                   # views: same list-of-tensors interface as the reference
            self.shadow_params.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.collected_params: List[torch.Tensor] = []
        self.steps = 0
        self._pending = 0                                   # fused updates applied but not yet acknowledged by update()

    # --- called by FusedAdamW.step ---------------------------------------------------------------
    def _next_decay(self) -> float:
        if self.num_updates is not None:
            self.num_updates += 1
        self._pending += 1
        return ema_decay_at(self.decay, self.num_updates)

    def update(self, parameters=None):
        """ema.py:36-53.  With a fused optimiser the update has already happened inside ``step()``."""
        if self._pending > 0:
            self._pending -= 1
            return
        raise _lib.BesoLibraryError("ExponentialMovingAverage.update(): attach the EMA to a FusedAdamW "
                                    "(opt.attach_ema(ema)); beso_b200 has no stand-alone / PyTorch EMA path")

    def copy_to(self, parameters):                          # ema.py:56-66
        for s, p in zip(self.shadow_params, [q for q in parameters if q.requires_grad]):
            p.data.copy_(s.data)

    def store(self, parameters):                            # ema.py:68-76
        self.collected_params = [p.clone() for p in parameters]

    def restore(self, parameters):                          # ema.py:78-89
        for c, p in zip(self.collected_params, parameters):
            p.data.copy_(c.data)

    def state_dict(self):
        return dict(decay=self.decay, num_updates=self.num_updates, shadow_params=self.shadow_params)

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        for s, v in zip(self.shadow_params, state_dict["shadow_params"]):
            s.copy_(v)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (single parameter group), one kernel launch per step."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        # config systems may hand over strings ("1e-4") or list-like betas
        super().__init__(params, dict(lr=float(lr), betas=(float(betas[0]), float(betas[1])), eps=float(eps),
                                      weight_decay=float(weight_decay)))
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group (the reference uses one)")
        self._params = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not self._params or not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() for p in self._params):
            raise _lib.BesoLibraryError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU path)")
        dev = self._params[0].device
        n = sum(p.numel() for p in self._params)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        self.step_count = 0
        self._ema: Optional[ExponentialMovingAverage] = None
        self._ema_every = 1
        self._handle = None
        self._ptrs = None

    def attach_ema(self, ema: ExponentialMovingAverage, update_every_n_steps: int = 1):
        """The EMA update of every ``update_every_n_steps``-th step is applied inside ``step()``
        (beso_agent.py:245-247)."""
        if ema.flat.numel() != self.exp_avg.numel() or ema.flat.device != self.exp_avg.device:
            raise ValueError("EMA and optimiser must cover the same parameters on the same device")
        self._ema, self._ema_every = ema, int(update_every_n_steps)

    def _opt(self):
        ptrs = tuple(p.data_ptr() for p in self._params)
        if self._handle is None or ptrs != self._ptrs:       # parameters were re-allocated (e.g. .to(device))
            self.close()
            arr = (C.c_void_p * len(ptrs))(*ptrs)
            numel = (C.c_longlong * len(ptrs))(*[p.numel() for p in self._params])
            h = C.c_void_p()
            _lib.check(_lib.lib().beso_opt_create(self._params[0].device.index or 0, len(ptrs), arr, numel, C.byref(h)),
                       "beso_opt_create")
            self._handle, self._ptrs = h, ptrs
        return self._handle

    @torch.no_grad()
    def step(self, closure=None, flat_grad: Optional[torch.Tensor] = None, grad_scale: float = 1.0):
        """``flat_grad``: the flat fp32 gradient in parameters() order (``model.last_flat_grad`` of the fused
        loss, or the all-reduced buffer); default = concatenation of ``p.grad``."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if flat_grad is None:
            if any(p.grad is None for p in self._params):
                raise _lib.BesoLibraryError("FusedAdamW.step(): every parameter needs a gradient")
            flat_grad = torch.cat([p.grad.reshape(-1) for p in self._params])
        if flat_grad.numel() != self.exp_avg.numel() or flat_grad.dtype != torch.float32 or not flat_grad.is_contiguous():
            raise ValueError("flat_grad must be a contiguous fp32 tensor covering all parameters")
        g = self.param_groups[0]
        self.step_count += 1
        ema_ptr, ema_decay = None, 0.0
        if self._ema is not None and self.step_count % self._ema_every == 0:
            ema_ptr, ema_decay = self._ema.flat.data_ptr(), self._ema._next_decay()
        stream = torch.cuda.current_stream(flat_grad.device).cuda_stream
        _lib.check(_lib.lib().beso_opt_step(self._opt(), flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                            ema_ptr, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                            float(g["weight_decay"]), self.step_count, float(ema_decay), float(grad_scale),
                                            C.c_void_p(stream)), "beso_opt_step")
        # the kernel wrote the parameters in place: bump their version counters like torch's in-place ops
        # would have, so that GCDenoiser re-packs its tensor-core weights (GCDenoiser._fingerprint)
        torch.autograd.graph.increment_version(self._params)
        return loss

    def close(self):
        if self._handle is not None:
            _lib.lib().beso_opt_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
