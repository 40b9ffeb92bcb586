"""Data scalers of the training / rollout path.

Mirrors of ``Scaler`` and ``MinMaxScaler`` (beso/networks/scaler/scaler_class.py:10-182, 185-374): same constructor,
attributes (``x_mean``, ``y_bounds``, ``y_bounds_tensor`` ...) and methods, statistics computed with numpy exactly as
there.  The per-call methods are a handful of elementwise ops on (B, T, dim) tensors and stay torch ops on the device;
the training path does not call them at all: ``gather_tables()`` hands the same arithmetic to the window-gather kernel
(beso_b200/csrc/dataset.cu), which applies it while it builds the batch.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

log = logging.getLogger(__name__)
_EPS = 1e-12
_PUSH_GOAL_FEATURES = [0, 1, 3, 4]   # block-push goals carry 4 of the 16 state features (scaler_class.py:158)


def _flatten(x_data, y_data):
    if isinstance(x_data, torch.Tensor):
        x_data, y_data = x_data.detach().cpu().numpy(), y_data.detach().cpu().numpy()
    if x_data.ndim == 3:      # (sequences, time, features) -> (samples, features)
        x_data, y_data = x_data.reshape(-1, x_data.shape[-1]), y_data.reshape(-1, y_data.shape[-1])
    elif x_data.ndim not in (2, 4):
        raise ValueError('not implemented yet!')
    return x_data, y_data


class _ScalerBase:
    def __init__(self, x_data, y_data, scale_data: bool, device):
        self.scale_data = scale_data
        self.device = device
        x_data, y_data = _flatten(x_data, y_data)
        dev = lambda a: torch.from_numpy(np.asarray(a)).to(device)  # noqa: E731
        self.x_mean, self.x_std = dev(x_data.mean(0)), dev(x_data.std(0))
        self.x_max, self.x_min = dev(x_data.max(0)), dev(x_data.min(0))
        self.y_min, self.y_max = dev(y_data.min(0)), dev(y_data.max(0))
        self._init_outputs(y_data, dev)
        self.x_bounds = np.zeros((2, x_data.shape[-1]))
        self.y_bounds = np.zeros((2, y_data.shape[-1]))
        if scale_data:
            x_den = x_data.std(0) + _EPS * np.ones(self.x_std.shape)
            self.x_bounds[0, :] = (x_data.min(0) - x_data.mean(0)) / x_den
            self.x_bounds[1, :] = (x_data.max(0) - x_data.mean(0)) / x_den
            self._scaled_output_bounds(y_data)
        else:
            self.x_bounds[0, :], self.x_bounds[1, :] = x_data.min(0), x_data.max(0)
            self.y_bounds[0, :], self.y_bounds[1, :] = y_data.min(0), y_data.max(0)
        self.y_bounds_tensor = torch.from_numpy(self.y_bounds).to(device)
        self.x_bounds_tensor = torch.from_numpy(self.x_bounds).to(device)
        self.tensor_y_bounds = torch.from_numpy(self.y_bounds).to(device)
        log.info('Datset Info: state min: {} and max: {}, action min: {} and max: {}'.format(
            self.x_bounds[0, :], self.x_bounds[1, :], self.y_bounds[0, :], self.y_bounds[1, :]))
        log.info(f'Training dataset size: input {x_data.shape} target {y_data.shape}')

    def _x_den(self, sel=None):
        std = self.x_std if sel is None else self.x_std[sel]
        return std + _EPS * torch.ones(std.shape, device=self.device)

    @torch.no_grad()
    def scale_input(self, x, block_push_goal=False):
        if x.shape[-1] == 4 and len(self.x_mean) == 16:     # block-push goal (scaler_class.py:88-90)
            return self.scale_block_push_goal(x)
        if x.shape[-1] == 7 and len(self.x_mean) == 30:     # one-hot kitchen goals are not scaled (scaler_class.py:92-93)
            return x.to(self.device)
        x = x.to(self.device)
        if not self.scale_data:
            return x
        return ((x - self.x_mean) / self._x_den()).to(torch.float32)

    @torch.no_grad()
    def scale_block_push_goal(self, x):
        x = x.to(self.device)
        if not self.scale_data:
            return x
        sel = _PUSH_GOAL_FEATURES   # the reference multiplies by x once more here (scaler_class.py:158); kept as is
        return x * (x - self.x_mean[sel]) / self._x_den(sel)

    @torch.no_grad()
    def clip_action(self, y):
        lo, hi = self.y_bounds_tensor[0, :] * 1.1, self.y_bounds_tensor[1, :] * 1.1
        return torch.clamp(y, lo, hi).to(self.device).to(torch.float32)

    # --- hand-over to the window-gather kernel -------------------------------------------------------------------
    def _tables(self):
        raise NotImplementedError

    def gather_tables(self):
        """(obs_table, act_table): (4, dim) fp32 device tensors with rows (sub, div, mul, add) such that
        ``((x - sub) / div) * mul + add`` is this scaler's scale_input / scale_output, or (None, None) when
        ``scale_data`` is off.  fp32 statistics only: with float64 data the reference computes in float64 and rounds
        at the end, which the fp32 kernel cannot reproduce bit for bit."""
        if not self.scale_data:
            return None, None
        if self.x_mean.dtype != torch.float32 or self.y_min.dtype != torch.float32:
            raise TypeError("fused scaling needs float32 statistics; build the scaler from float32 data")
        if getattr(self, "_gather_tables", None) is None:
            self._gather_tables = tuple(torch.stack(rows).contiguous() for rows in self._tables())
        return self._gather_tables


    def _inverse_output_rows(self):
        raise NotImplementedError

    def rollout_tables(self):
        """(in_table, out_table, clip) for the sampling kernel's fused rollout scaling (``beso_io_scaling``):
        ``scale_input`` and ``inverse_scale_output`` as (4, dim) fp32 tables with rows (sub, div, mul, add) -- None when
        ``scale_data`` is off -- and the float64 (2, act) bounds of ``clip_action``."""
        clip = torch.stack([self.y_bounds_tensor[0, :] * 1.1, self.y_bounds_tensor[1, :] * 1.1]).to(torch.float64).contiguous()
        if not self.scale_data:
            return None, None, clip
        in_table, _ = self.gather_tables()
        if getattr(self, "_inverse_table", None) is None:
            self._inverse_table = torch.stack(self._inverse_output_rows()).to(torch.float32).contiguous()
        return in_table, self._inverse_table, clip


class Scaler(_ScalerBase):
    """Standardises inputs and outputs with the data's mean and standard deviation (scaler_class.py:10-182)."""

    def _init_outputs(self, y_data, dev):
        self.y_mean, self.y_std = dev(y_data.mean(0)), dev(y_data.std(0))

    def _scaled_output_bounds(self, y_data):
        y_den = y_data.std(0) + _EPS * np.ones(self.y_std.shape)
        self.y_bounds[0, :] = (y_data.min(0) - y_data.mean(0)) / y_den
        self.y_bounds[1, :] = (y_data.max(0) - y_data.mean(0)) / y_den

    def _y_den(self):
        return self.y_std + _EPS * torch.ones(self.y_std.shape, device=self.device)

    @torch.no_grad()
    def scale_output(self, y):
        y = y.to(self.device)
        return ((y - self.y_mean) / self._y_den()).to(torch.float32) if self.scale_data else y

    @torch.no_grad()
    def inverse_scale_input(self, x):
        return (x * self._x_den() + self.x_mean).to(torch.float32) if self.scale_data else x.to(self.device)

    @torch.no_grad()
    def inverse_scale_output(self, y):
        return y * self._y_den() + self.y_mean if self.scale_data else y.to(self.device)

    def _tables(self):
        one, zero = torch.ones_like, torch.zeros_like
        return ([self.x_mean, self._x_den(), one(self.x_mean), zero(self.x_mean)],
                [self.y_mean, self._y_den(), one(self.y_mean), zero(self.y_mean)])

    def _inverse_output_rows(self):          # y * den + mean
        one, zero = torch.ones_like, torch.zeros_like
        return [zero(self.y_mean), one(self.y_mean), self._y_den(), self.y_mean]


class MinMaxScaler(_ScalerBase):
    """Inputs standardised, outputs mapped linearly from [y_min, y_max] to [-1, 1] (scaler_class.py:185-374)."""

    def _init_outputs(self, y_data, dev):
        self.new_max_x, self.new_min_x = torch.ones_like(self.x_max), -1 * torch.ones_like(self.x_max)
        self.new_max_y, self.new_min_y = torch.ones_like(self.y_max), -1 * torch.ones_like(self.y_max)

    def _scaled_output_bounds(self, y_data):
        self.y_bounds[0, :], self.y_bounds[1, :] = -1.0, 1.0

    @torch.no_grad()
    def scale_output(self, y):
        y = y.to(self.device)
        if not self.scale_data:
            return y
        out = (y - self.y_min) / (self.y_max - self.y_min) * (self.new_max_y - self.new_min_y) + self.new_min_y
        return out.to(torch.float32)

    @torch.no_grad()
    def inverse_scale_input(self, x):
        if not self.scale_data:
            return x.to(self.device)
        out = (x - self.new_min_x) / (self.new_max_x - self.new_min_x) * (self.x_max - self.x_min) + self.x_min
        return out.to(torch.float32)

    @torch.no_grad()
    def inverse_scale_output(self, y):
        if not self.scale_data:
            return y.to(self.device)
        return (y - self.new_min_y) / (self.new_max_y - self.new_min_y) * (self.y_max - self.y_min) + self.y_min

    def _tables(self):
        one, zero = torch.ones_like, torch.zeros_like
        return ([self.x_mean, self._x_den(), one(self.x_mean), zero(self.x_mean)],
                [self.y_min, self.y_max - self.y_min, self.new_max_y - self.new_min_y, self.new_min_y])

    def _inverse_output_rows(self):          # (y - new_min) / (new_max - new_min) * (y_max - y_min) + y_min
        return [self.new_min_y, self.new_max_y - self.new_min_y, self.y_max - self.y_min, self.y_min]
