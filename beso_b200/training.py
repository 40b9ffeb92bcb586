"""Training path: ``GCDenoiser.loss`` (k_diffusion/score_wrappers.py:45-79) with a hand-derived backward.

``denoiser_loss`` returns a scalar tensor wired into autograd: ``loss.backward()`` fills ``.grad`` of every
parameter exactly like the reference's autograd would (beso_agent.py:236-240).  Forward and backward run
in ONE call of the C ABI (``beso_loss_fwd_bwd``), which writes all gradients into one flat fp32 buffer in
``parameters()`` order -- the buffer the data-parallel all-reduce (beso_b200/dist.py) operates on.

Every dense product runs on the tcgen05 tensor cores (csrc/gemm.cu): ``model.train_math = "fp32"`` (default) splits
the operands into three bf16 images, six MMAs per product (the fp32-parity mode the gradient goldens pin);
``"bf16x2"`` is two images / three MMAs, ``"bf16"`` one MMA per product.

Training-mode randomness: the reference draws the element-wise goal mask of CFG training (score_gpts.py:360-371) and
its dropout masks (score_gpts.py:338,72,79,109) from torch's global generator in op order.
``draw_dropout_masks`` makes the same torch calls in the same order here and the kernels apply the masks, so a
training step consumes the generator exactly like the reference's and -- on the same device type -- sees the same
masks.
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


def draw_dropout_masks(inner, batch: int, t: int, device):
    """The nn.Dropout draws of one training forward of ``DiffusionGPT`` in the reference's op order:
    ``self.drop(input_seq)`` (score_gpts.py:338), then per block ``attn_drop`` (:72, shape (B, H, T, T)),
    ``resid_drop`` (:79) and the mlp's Dropout (:109), both (B, T, d).  Each mask is what F.dropout applies to a tensor
    of ones (0 or 1 / (1 - p)); F.dropout consumes the generator by shape only, so the global generator advances
    exactly as in the reference.  p = 0 draws nothing (as in torch).  Returns None when every p is 0."""
    p_embed, p_attn, p_resid = inner._dropouts
    if not inner.training or not any(p > 0 for p in (p_embed, p_attn, p_resid)):
        return None
    cfg = inner.config
    T, d, H, L = cfg.n_tokens(t), cfg.d, cfg.n_heads, cfg.n_layers

    def draw(shape, p):
        if p <= 0:
            return None
        return torch.nn.functional.dropout(torch.ones(shape, device=device, dtype=torch.float32), p, True)

    masks = {"embed": draw((batch, T, d), p_embed), "attn": [], "resid_attn": [], "resid_mlp": []}
    for _ in range(L):
        masks["attn"].append(draw((batch, H, T, T), p_attn))
        masks["resid_attn"].append(draw((batch, T, d), p_resid))
        masks["resid_mlp"].append(draw((batch, T, d), p_resid))
    return masks


def _masks_struct(masks, n_layers, device):
    """ctypes view of ``masks`` (beso_dropout_masks); returns (struct or None, keep-alive list)."""
    if masks is None:
        return None, []
    keep = []

    def ptr(t):
        if t is None:
            return None
        t = t.to(device=device, dtype=torch.float32).contiguous()
        keep.append(t)
        return t.data_ptr()

    def arr(lst):
        if lst is None or all(m is None for m in lst):
            return None
        if len(lst) != n_layers or any(m is None for m in lst):
            raise ValueError("dropout masks: need one mask per layer")
        a = (C.c_void_p * n_layers)(*[ptr(m) for m in lst])
        keep.append(a)
        return a

    s = _lib.DropoutMasks()
    s.embed = ptr(masks.get("embed"))
    for name in ("attn", "resid_attn", "resid_mlp"):
        a = arr(masks.get(name))
        if a is not None:
            setattr(s, name, C.cast(a, C.POINTER(C.c_void_p)))
    return s, keep


def loss_and_flat_grad(model, state, action, goal, noise, sigma, pred_last_action_only=False, goal_keep=None,
                       need_grad=True, dropout_masks=None, grad_sync=None):
    """Runs ``beso_loss_fwd_bwd_dropout``; returns (loss 0-d tensor, flat gradient or None).

    ``model.train_math``: "fp32" (default: three bf16 images per operand, six MMAs per product, the fp32-parity mode),
    "bf16x2" (two images, three MMAs) or "bf16" (one MMA per product; "tf32" is accepted as the round-1 name of the
    opt-in fast mode).
    ``dropout_masks``: see ``draw_dropout_masks`` (None = no dropout).
    ``grad_sync``: a ``beso_b200.dist.FlatGradAllReduce``; with the NCCL transport the gradient all-reduce is issued per
    transformer block from inside the backward pass on the communicator's stream (``beso_loss_fwd_bwd_dp``) and the
    returned flat gradient is already the mean over the ranks; other transports reduce it after the call."""
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
    if math not in ("fp32", "bf16x2", "bf16", "tf32"):
        raise ValueError(f"train_math must be 'fp32', 'bf16x2' or 'bf16', got {math!r}")
    if math == "bf16x2":
        flags |= _lib.FLAG_TRAIN_SPLIT2
    elif math != "fp32":
        flags |= _lib.FLAG_TRAIN_FAST
    mstruct, keep = _masks_struct(dropout_masks, cfg.n_layers, action.device)
    stream = torch.cuda.current_stream(dev).cuda_stream
    comm = grad_sync.comm_handle() if (grad_sync is not None and flat is not None) else None
    _lib.check(_lib.lib().beso_loss_fwd_bwd_dp(plan, state.data_ptr(), action.data_ptr(), goal.data_ptr(),
                                              noise.data_ptr(), sigma.data_ptr(), keep_ptr,
                                              C.byref(mstruct) if mstruct is not None else None, comm,
                                              1.0 / grad_sync.world if comm is not None else 1.0, loss.data_ptr(),
                                              flat.data_ptr() if flat is not None else None, B, flags,
                                              C.c_void_p(stream)),
               "beso_loss_fwd_bwd_dp")
    del keep
    if grad_sync is not None and comm is None and flat is not None:
        grad_sync(flat)                                     # torch.distributed transport (or world 1): after the call
    return loss, flat


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, state, action, goal, noise, sigma, pred_last, goal_keep, drop, *params):
        need = any(p.requires_grad for p in params) and torch.is_grad_enabled()
        loss, flat = loss_and_flat_grad(model, state, action, goal, noise, sigma, pred_last, goal_keep, need_grad=True,
                                        dropout_masks=drop)
        ctx.model = model
        ctx.flat = flat
        model.last_flat_grad = flat            # exposed for the data-parallel exchange
        del need
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        views = flat_grad_views(ctx.model, ctx.flat)
        grads = tuple(v * grad_out for v in views)
        return (None,) * 9 + grads


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
    drop = draw_dropout_masks(inner, action.shape[0], action.shape[1], action.device)   # after mask_cond, in op order
    params = _param_list(model)
    return _LossFn.apply(model, state, action, goal, noise, sigma, pred_last, goal_keep, drop, *params)
