"""Training path: GCDenoiser.loss (score_wrappers.py:45-79) with hand-written backward kernels."""
from __future__ import annotations

from . import _lib


def denoiser_loss(model, state, action, goal, noise, sigma, **kwargs):
    raise _lib.BesoLibraryError("beso_loss_fwd_bwd is not implemented in this build")
