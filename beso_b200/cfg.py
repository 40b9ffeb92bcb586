"""Classifier-free-guidance sampling wrapper (k_diffusion/classifier_free_sampler.py:12-52).

``out_uncond + cond_lambda * (out_cond - out_uncond)``; with a beso_b200 ``GCDenoiser`` inside,
both branches and the mix run in ONE kernel launch (BESO_FLAG_CFG).
"""
from __future__ import annotations

import torch.nn as nn

from .denoiser import GCDenoiser


class ClassifierFreeSampleModel(nn.Module):
    def __init__(self, model, cond_lambda: float = 2):
        super().__init__()
        self.model = model
        self.cond_lambda = cond_lambda
        self.cond = cond_lambda == 1

    def forward(self, state, action, goal, sigma, **extra_args):
        if self.cond:                                  # lambda == 1: conditional branch only
            return self.model(state, action, goal, sigma)
        if self.cond_lambda == 0:                      # lambda == 0: unconditional branch only
            return self.model(state, action, goal, sigma, uncond=True)
        if isinstance(self.model, GCDenoiser) and not extra_args:
            return self.model._run(state, action, goal, sigma, cfg_lambda=float(self.cond_lambda))
        out = self.model(state, action, goal, sigma, **extra_args)
        out_uncond = self.model(state, action, goal, sigma, uncond=True)
        return out_uncond + self.cond_lambda * (out - out_uncond)

    def get_params(self):
        return self.model.get_params()
