"""Training path: ``GCDenoiser.loss`` (k_diffusion/score_wrappers.py:45-79) with a hand-derived backward.

``denoiser_loss`` returns a scalar tensor wired into autograd: ``loss.backward()`` fills ``.grad`` of every
parameter exactly like the reference's autograd would (beso_agent.py:236-240).  Forward and backward run
in ONE call of the C ABI (``beso_loss_fwd_bwd``), which writes all gradients into one flat fp32 buffer in
``parameters()`` order -- the buffer the data-parallel all-reduce (beso_b200/dist.py) operates on.

Dropout must be off (p = 0): the reference draws dropout masks from the global RNG in op order
(SURVEY.md H5).  The element-wise goal mask of CFG training (score_gpts.py:360-371) is drawn here with
the same torch call and passed to the kernel.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _param_list(model):
    return list(model.inner_model.parameters())


def flat_grad_views(model, flat: torch.Tensor):
    """Views of the flat gradient buffer, one per parameter, in parameters() order."""
    out, off = [], 0
    for p in _param_list(model):
        out.append(flat[off:off + p.numel()].view_as(p))
        off += p.numel()
    return out


def loss_and_flat_grad(model, state, action, goal, noise, sigma, pred_last_action_only=False, goal_keep=None,
                       need_grad=True):
    """Runs ``beso_loss_fwd_bwd``; returns (loss 0-d tensor, flat gradient or None).

    ``model.train_math = "tf32"`` (default "fp32", the reference's arithmetic) runs the GEMMs on the tensor
    cores in TF32."""
    inner = model.inner_model
    if any(p > 0 for p in inner._dropouts) and inner.training:
        raise _lib.BesoLibraryError("the fused training path needs attn_pdrop = resid_pdrop = embed_pdrob = 0")
    params = _param_list(model)
    dev = model._device_index(action)
    state, action, goal, noise, sigma = map(model._prep, (state, action, goal, noise, sigma))
    B, t = model._check_shapes(state, action, goal)
    cfg = model.config
    if t != cfg.window:
        raise ValueError(f"training needs full windows (t = {cfg.window}), got t = {t}")
    plan = model._ensure_plan(dev)
    ptrs = (C.c_void_p * len(params))(*[p.data_ptr() for p in params])
    _lib.check(_lib.lib().beso_plan_set_params(plan, model._slot, ptrs, len(params)), "beso_plan_set_params")
    loss = torch.empty((), device=action.device, dtype=torch.float32)
    flat = torch.empty(sum(p.numel() for p in params), device=action.device, dtype=torch.float32) if need_grad else None
    keep_ptr = goal_keep.data_ptr() if goal_keep is not None else None
    flags = _lib.FLAG_PRED_LAST if pred_last_action_only else 0
    math = getattr(model, "train_math", "fp32")
    if math not in ("fp32", "tf32"):
        raise ValueError(f"train_math must be 'fp32' or 'tf32', got {math!r}")
    if math == "tf32":
        flags |= _lib.FLAG_TRAIN_TF32
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(_lib.lib().beso_loss_fwd_bwd(plan, state.data_ptr(), action.data_ptr(), goal.data_ptr(), noise.data_ptr(),
                                           sigma.data_ptr(), keep_ptr, loss.data_ptr(),
                                           flat.data_ptr() if flat is not None else None, B, flags, C.c_void_p(stream)),
               "beso_loss_fwd_bwd")
    return loss, flat


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, state, action, goal, noise, sigma, pred_last, goal_keep, *params):
        need = any(p.requires_grad for p in params) and torch.is_grad_enabled()
        loss, flat = loss_and_flat_grad(model, state, action, goal, noise, sigma, pred_last, goal_keep, need_grad=True)
        ctx.model = model
        ctx.flat = flat
        model.last_flat_grad = flat            # exposed for the data-parallel exchange
        del need
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        views = flat_grad_views(ctx.model, ctx.flat)
        grads = tuple(v * grad_out for v in views)
        return (None,) * 8 + grads


def denoiser_loss(model, state, action, goal, noise, sigma, **kwargs):
    """GCDenoiser.loss(state, action, goal, noise, sigma, pred_last_action_only=False)."""
    pred_last = bool(kwargs.pop("pred_last_action_only", False))
    if kwargs:
        raise TypeError(f"unexpected keyword arguments {sorted(kwargs)}")
    inner = model.inner_model
    if pred_last:
        noise[:, :-1, :] = 0                   # the reference mutates the caller's noise (score_wrappers.py:63)
    goal_keep = None
    if inner.training and inner.cond_mask_prob > 0.0:
        mask = torch.bernoulli(torch.ones(goal.shape, device=goal.device) * inner.cond_mask_prob)
        goal_keep = (1.0 - mask).contiguous()
    params = _param_list(model)
    return _LossFn.apply(model, state, action, goal, noise, sigma, pred_last, goal_keep, *params)
