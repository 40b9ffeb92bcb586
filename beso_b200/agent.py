"""Host-side mirror of the hot-path half of ``BesoAgent`` (diffusion_agents/beso_agent.py).

Only what sits directly on the hot path is here: ``sample_loop`` (beso_agent.py:390-456),
``get_noise_schedule`` (:580-598), ``evaluate`` (:251-289), the stateful ``predict``
(:297-388) with its observation / action context deques, and ``train_step`` (:215-248) with
``make_sample_density`` (:540-578) on top of the fused loss / backward and the fused AdamW + EMA
step.  Hydra, wandb, workspaces and data loading stay in the reference shell; to use the
reference's own BesoAgent instead, point its model ``_target_`` at
``beso_b200.denoiser.GCDenoiser``.
"""
from __future__ import annotations

import math
from collections import deque
from typing import Optional

import torch

from . import sampling
from .cfg import ClassifierFreeSampleModel
from .denoiser import GCDenoiser

FUSED_SAMPLERS = ("ddim", "euler", "heun")
P = "inner_model."   # state_dict prefix of the score network inside GCDenoiser


class IdentityScaler:
    """Stands in for networks/scaler/scaler_class.py when data is not scaled (scale_data: False)."""

    def scale_input(self, x):
        return x

    def scale_output(self, x):
        return x

    def inverse_scale_output(self, x):
        return x

    def clip_action(self, x):
        return x


class BesoAgent:
    def __init__(self, model: GCDenoiser, device="cuda", sampler_type: str = "ddim", num_sampling_steps: int = 3,
                 sigma_min: float = 0.005, sigma_max: float = 1.0, rho: float = 5.0, window_size: int = 4,
                 noise_scheduler: str = "exponential", use_ema: bool = False, ema_params=None,
                 scaler=None, pred_last_action_only: bool = False, cond_lambda: Optional[float] = None):
        self.model = model
        self.device = device
        self.sampler_type = sampler_type
        self.num_sampling_steps = num_sampling_steps
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho
        self.window_size = window_size
        self.noise_scheduler = noise_scheduler
        self.use_ema = use_ema
        self.ema_params = ema_params          # list of tensors in parameters() order, or None
        self.scaler = scaler if scaler is not None else IdentityScaler()
        self.pred_last_action_only = pred_last_action_only
        self.obs_context = deque(maxlen=window_size)
        self.action_context = deque(maxlen=window_size - 1)
        if cond_lambda is not None:
            self.model = ClassifierFreeSampleModel(model, cond_lambda)

    # ---- base_agent.py:111-142 --------------------------------------------------------------
    GOAL_ZERO_DIMS = [2, 5, 6, 7, 8, 9]   # block-push goals (10 features): only the block positions are a goal

    def process_batch(self, batch: dict, predict: bool = True):
        """Scaled (state, action, goal) of a batch dict.  Batches from ``DeviceWindowDataset`` with a scaler attached
        are already scaled (``"scaled": True``).  As in the reference, 10-feature goals get dims [2, 5, 6, 7, 8, 9]
        zeroed after scaling, and ``predict=True`` without an action returns ``(state, goal, goal_task_name | None)``."""
        pre = batch.get("scaled", False)
        state, goal = batch["observation"].to(self.device), batch["goal_observation"].to(self.device)
        if not pre:
            state, goal = self.scaler.scale_input(state), self.scaler.scale_input(goal)
        if goal.shape[-1] == 10:
            goal[..., self.GOAL_ZERO_DIMS] = 0
        if "action" in batch:
            action = batch["action"].to(self.device)
            return state, (action if pre else self.scaler.scale_output(action)), goal
        if predict:
            return state, goal, batch.get("goal_task_name")
        return state, goal

    def _rollout_scaling(self, batch: dict):
        """``RolloutScaling`` for this batch when its scaler calls can ride inside the sampling kernel:
        a scaler with ``rollout_tables`` (beso_b200.scaler), fp32 statistics, states and goals with the model's
        feature count (the 4-feature block-push goals and one-hot kitchen goals of scaler_class.py:88-93 keep the torch
        path), batch not pre-scaled.  None otherwise."""
        from .denoiser import RolloutScaling
        if batch.get("scaled", False) or not hasattr(self.scaler, "rollout_tables"):
            return None
        obs = self._core().config.obs_dim
        if batch["observation"].shape[-1] != obs or batch["goal_observation"].shape[-1] != obs:
            return None
        try:
            in_table, out_table, clip = self.scaler.rollout_tables()
        except TypeError:                                    # float64 statistics: the fp32 kernel cannot replay them
            return None
        if in_table is not None and in_table.shape[-1] != obs:
            return None
        goal_keep = None
        if obs == 10:                                        # base_agent.py:119-120
            goal_keep = torch.ones(obs, device=self.device, dtype=torch.float32)
            goal_keep[self.GOAL_ZERO_DIMS] = 0
        dev = torch.device(self.device)
        mv = lambda t: None if t is None else t.to(dev).contiguous()   # noqa: E731
        return RolloutScaling(mv(in_table), goal_keep, mv(clip), mv(out_table))

    # ---- beso_agent.py:106-117, 458-477: scaler hand-over and the checkpoint files ------------------
    def get_scaler(self, scaler):
        self.scaler = scaler

    def set_bounds(self, scaler):
        self.model.min_action = torch.from_numpy(scaler.y_bounds[0, :]).to(self.device)
        self.model.max_action = torch.from_numpy(scaler.y_bounds[1, :]).to(self.device)

    def load_pretrained_model(self, weights_path: str, **kwargs) -> None:
        """``<weights_path>/model_state_dict.pth`` (the EMA weights the reference ships) into the model; the EMA helper
        restarts from them.  ``map_location`` is set (the shipped files hold CUDA storages)."""
        import os
        sd = torch.load(os.path.join(weights_path, "model_state_dict.pth"), map_location="cpu")
        core = self._core()
        core.load_state_dict(sd)
        self.ema_updated()
        if getattr(self, "ema_helper", None) is not None:
            from .optim import ExponentialMovingAverage
            self.ema_helper = ExponentialMovingAverage(list(core.get_params()), self.ema_helper.decay, self.device)
            self.ema_params = self.ema_helper.shadow_params
            if getattr(self, "optimizer", None) is not None:
                self.optimizer.attach_ema(self.ema_helper, self.update_ema_every_n_steps)

    def store_model_weights(self, store_path: str) -> None:
        """The reference's two files: ``model_state_dict.pth`` with the EMA weights in place of the parameters (when
        EMA is in use) and ``non_ema_model_state_dict.pth`` with the raw ones.  Nothing is swapped in the live model."""
        import os
        core = self._core()
        raw = {k: v.detach().clone() for k, v in core.state_dict().items()}
        ema = dict(raw)
        if self.use_ema and self.ema_params is not None:
            names = [P + n for n, _ in core.inner_model.named_parameters()]
            for n, e in zip(names, self.ema_params):
                ema[n] = e.detach().clone().view_as(raw[n])
        torch.save(ema, os.path.join(store_path, "model_state_dict.pth"))
        torch.save(raw, os.path.join(store_path, "non_ema_model_state_dict.pth"))

    # ---- beso_agent.py:580-598 --------------------------------------------------------------
    def get_noise_schedule(self, n_sampling_steps, noise_schedule_type):
        s = sampling
        if noise_schedule_type == "karras":
            return s.get_sigmas_karras(n_sampling_steps, self.sigma_min, self.sigma_max, self.rho, self.device)
        if noise_schedule_type == "exponential":
            return s.get_sigmas_exponential(n_sampling_steps, self.sigma_min, self.sigma_max, self.device)
        if noise_schedule_type == "vp":
            return s.get_sigmas_vp(n_sampling_steps, device=self.device)
        if noise_schedule_type == "linear":
            return s.get_sigmas_linear(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        if noise_schedule_type == "ve":
            return s.get_sigmas_ve(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        raise ValueError("Unknown noise schedule type")

    # ---- beso_agent.py:390-456 --------------------------------------------------------------
    def sample_loop(self, sigmas, x_t, state, goal, sampler_type, extra_args={}):
        s_churn = extra_args["s_churn"] if "s_churn" in extra_args else 0
        s_min = extra_args["s_min"] if "s_min" in extra_args else 0
        use_scaler = extra_args["use_scaler"] if "use_scaler" in extra_args else False
        if bool(extra_args):
            _ = {k: extra_args[k] for k in ("s_churn", "keep_last_actions")}   # KeyError like the reference
        scaler = self.scaler if use_scaler else None
        if sampler_type == "heun":
            return sampling.sample_heun(self.model, state, x_t, goal, sigmas, scaler=scaler, s_churn=s_churn,
                                        s_tmin=s_min, disable=True)
        if sampler_type == "euler":
            return sampling.sample_euler(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "ddim":
            return sampling.sample_ddim(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "lms":                          # beso_agent.py:419-420
            return sampling.sample_lms(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "dpm":                          # beso_agent.py:434-435 (dispatched without the scaler)
            return sampling.sample_dpm_2(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == "ancestral":                    # beso_agent.py:428-429
            return sampling.sample_dpm_2_ancestral(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "dpmpp_2s_ancestral":           # beso_agent.py:445-446
            return sampling.sample_dpmpp_2s_ancestral(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "dpmpp_2s":                     # beso_agent.py:447-448
            return sampling.sample_dpmpp_2s(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "dpmpp_2m":                     # beso_agent.py:450-451
            return sampling.sample_dpmpp_2m(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "euler_ancestral":              # beso_agent.py:431-432
            return sampling.sample_euler_ancestral(self.model, state, x_t, goal, sigmas, scaler=scaler, disable=True)
        if sampler_type == "dpm_adaptive":                 # beso_agent.py:438-439 (step control on the host)
            return sampling.sample_dpm_adaptive(self.model, state, x_t, goal, sigmas[-2].item(), sigmas[0].item(), disable=True)
        if sampler_type == "dpm_fast":                     # beso_agent.py:441-442
            return sampling.sample_dpm_fast(self.model, state, x_t, goal, sigmas[-2].item(), sigmas[0].item(), len(sigmas),
                                            disable=True)
        raise ValueError("desired sampler type not found!")

    def _core(self) -> GCDenoiser:
        return self.model.model if isinstance(self.model, ClassifierFreeSampleModel) else self.model

    def _use_ema_weights(self, on: bool):
        """store/copy_to/restore of the reference EMA helper (beso_agent.py:343-345,380-381) without
        moving a byte: the EMA copy is packed once into weight slot 1 and selected by pointer."""
        if not (self.use_ema and self.ema_params is not None):
            return
        core = self._core()
        if on:
            # Slot 1 is packed straight from the EMA shadow tensors and stamped with the EMA generation it was packed
            # at; optimiser steps that do not update the EMA (update_ema_every_n_steps > 1) change neither.
            gen = getattr(self, "_ema_generation", 0)
            if 1 not in core._packed or core.packed_tag(1) != ("ema", gen):
                core.refresh_weights(slot=0, force=False)          # plan exists, raw slot current
                core.pack_tensors(1, list(self.ema_params), tag=("ema", gen))
            core.select_weights(1)
        else:
            core.select_weights(0)

    def ema_updated(self):
        """Call after the EMA shadow parameters changed: slot 1 is re-packed from them on next use."""
        self._ema_generation = getattr(self, "_ema_generation", 0) + 1

    # ---- training: beso_agent.py:215-248, 540-578; k_diffusion/utils.py:170-200 -------------------------
    def configure_training(self, lr: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 1e-2, eps: float = 1e-8,
                           lr_step_size: int = 100, lr_gamma: float = 0.99, decay: float = 0.999,
                           update_ema_every_n_steps: int = 1, sigma_sample_density_type: str = "loglogistic",
                           sigma_sample_density_mean: float = -1.2, sigma_sample_density_std: float = 1.2):
        """Optimiser, StepLR and EMA of the reference agent (configs/agents/beso_kitchen.yaml:9-17, 26-31;
        base_agent.py:33-40, beso_agent.py:60-66), with the fused AdamW + EMA step."""
        from .optim import ExponentialMovingAverage, FusedAdamW
        core = self._core()
        params = list(core.get_params())
        self.optimizer = FusedAdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.lr_scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=lr_step_size, gamma=lr_gamma)
        self.ema_helper = ExponentialMovingAverage(params, decay, self.device)
        self.update_ema_every_n_steps = update_ema_every_n_steps
        self.optimizer.attach_ema(self.ema_helper, update_ema_every_n_steps)
        self.use_ema, self.ema_params = True, self.ema_helper.shadow_params
        self.sigma_sample_density_type = sigma_sample_density_type
        self.sigma_sample_density_mean, self.sigma_sample_density_std = sigma_sample_density_mean, sigma_sample_density_std
        self.steps = 0
        self.grad_sync = getattr(self, "grad_sync", None)

    def enable_data_parallel(self, transport: str = "nccl"):
        """Data-parallel training (BASELINE config 4): every rank runs ``train_step`` on its own shard of
        the global batch (``DeviceWindowDataset.batches(rank=..., world_size=...)``); the flat gradient is summed over
        the ranks with ONE all-reduce and scaled by 1/world before the optimiser step, so replicas that start equal stay
        bit-identical.  ``torch.distributed`` must be initialised; ``transport`` as in ``beso_b200.dist``."""
        from .dist import FlatGradAllReduce
        index = None
        if transport == "nccl":
            dev = torch.device(self.device)
            index = dev.index if dev.index is not None else torch.cuda.current_device()
        self.grad_sync = FlatGradAllReduce(transport, device=index)
        return self.grad_sync

    def make_sample_density(self):
        """Noise-level distribution for training (beso_agent.py:540-578; the four types the shipped configs use)."""
        sigma_data = self._core().sigma_data
        kind = self.sigma_sample_density_type
        if kind == "lognormal":
            loc, scale = self.sigma_sample_density_mean, self.sigma_sample_density_std
            return lambda shape, device: (torch.randn(shape, device=device, dtype=torch.float32) * scale + loc).exp()
        if kind == "loglogistic":                           # utils.rand_log_logistic, truncated to [sigma_min, sigma_max]
            loc, scale = math.log(sigma_data), 0.5

            def draw(shape, device):
                lo = torch.as_tensor(self.sigma_min, device=device, dtype=torch.float64)
                hi = torch.as_tensor(self.sigma_max, device=device, dtype=torch.float64)
                min_cdf, max_cdf = lo.log().sub(loc).div(scale).sigmoid(), hi.log().sub(loc).div(scale).sigmoid()
                u = torch.rand(shape, device=device, dtype=torch.float64) * (max_cdf - min_cdf) + min_cdf
                return u.logit().mul(scale).add(loc).exp().to(torch.float32)
            return draw
        if kind == "loguniform":
            lo, hi = math.log(self.sigma_min), math.log(self.sigma_max)
            return lambda shape, device: (torch.rand(shape, device=device, dtype=torch.float32) * (hi - lo) + lo).exp()
        if kind == "uniform":
            lo, hi = self.sigma_min, self.sigma_max
            return lambda shape, device: torch.rand(shape, device=device, dtype=torch.float32) * (hi - lo) + lo
        raise ValueError("Unknown sample density type")

    def train_step(self, batch: dict) -> float:
        """One optimisation step (beso_agent.py:215-248): noise and noise levels are drawn with the same torch calls,
        loss + all gradients come from ONE call of the fused forward / backward, AdamW + EMA are ONE launch."""
        from .training import loss_and_flat_grad
        core = self._core()
        state, action, goal = self.process_batch(batch, predict=False)
        core.train()
        core.training = True
        noise = torch.randn_like(action)
        sigma = self.make_sample_density()(shape=(len(action),), device=self.device)
        inner = core.inner_model
        goal_keep = None
        if inner.cond_mask_prob > 0.0:                      # element-wise goal mask of CFG training (score_gpts.py:360-371)
            mask = torch.bernoulli(torch.ones(goal.shape, device=goal.device) * inner.cond_mask_prob)
            goal_keep = (1.0 - mask).contiguous()
        if self.pred_last_action_only:
            noise[:, :-1, :] = 0
        from .training import draw_dropout_masks
        drop = draw_dropout_masks(inner, action.shape[0], action.shape[1], action.device)   # nn.Dropout draws, in op order
        # data parallel: the mean of the ranks' gradients, all-reduced per block from inside the backward pass
        gs = getattr(self, "grad_sync", None)
        overlapped = gs is not None and getattr(gs, "comm_handle", lambda: None)() is not None
        loss, flat = loss_and_flat_grad(core, state, action, goal, noise, sigma, self.pred_last_action_only, goal_keep,
                                        dropout_masks=drop, grad_sync=gs if overlapped else None)
        if gs is not None and not overlapped:               # torch.distributed transport: one all-reduce after the call
            gs(flat)
        self.optimizer.step(flat_grad=flat)                 # zero_grad / backward / step of the reference in one
        self.lr_scheduler.step()
        self.steps += 1
        if self.steps % self.update_ema_every_n_steps == 0:
            self.ema_helper.update(core.parameters())       # acknowledged: applied inside optimizer.step
            self.ema_updated()
        return loss.item()

    # ---- beso_agent.py:251-289 --------------------------------------------------------------
    @torch.no_grad()
    def evaluate(self, batch, action=None, goal=None) -> float:
        """``evaluate(batch)`` as the reference's workspaces call it (beso_agent.py:251-289: the batch dict goes through
        ``process_batch``), or ``evaluate(state, action, goal)`` with already scaled tensors."""
        if isinstance(batch, dict):
            state, action, goal = self.process_batch(batch, predict=True)
        else:
            state = batch
        self._use_ema_weights(True)
        self._core().eval()
        sigmas = sampling.get_sigmas_exponential(self.num_sampling_steps, self.sigma_min, self.sigma_max, self.device)
        x = torch.randn_like(action) * self.sigma_max
        x_0 = self.sample_loop(sigmas, x, state, goal, self.sampler_type)
        if self.pred_last_action_only and x_0.dim() == 2:
            x_0 = x_0.unsqueeze(1)
        mse = torch.nn.functional.mse_loss(x_0, action, reduction="none").mean().item()
        self._use_ema_weights(False)
        return mse

    def reset(self):
        self.obs_context.clear()
        self.action_context.clear()

    # ---- beso_agent.py:297-388 --------------------------------------------------------------
    @torch.no_grad()
    def predict(self, batch: dict, new_sampler_type=None, get_mean=None, new_sampling_steps=None,
                extra_args=None, noise_scheduler=None) -> torch.Tensor:
        extra_args = {} if extra_args is None else extra_args
        # scaler prologue / epilogue inside the sampling kernel when the scaler allows it; decided once per episode so
        # that the observation context holds either raw or scaled states, never a mix
        if len(self.obs_context) == 0:
            self._fused_io = self._rollout_scaling(batch)
        io = getattr(self, "_fused_io", None)
        if io is not None:
            state, goal = batch["observation"].to(self.device), batch["goal_observation"].to(self.device)   # raw
        else:
            state, goal, _ = self.process_batch(batch, predict=True)
        n_steps = new_sampling_steps if new_sampling_steps is not None else self.num_sampling_steps
        sampler_type = new_sampler_type if new_sampler_type is not None else self.sampler_type
        self.obs_context.append(state)
        input_state = torch.stack(tuple(self.obs_context), dim=1)
        if goal.dim() == 2:
            goal = goal.unsqueeze(0)
        self._use_ema_weights(True)
        self._core().eval()
        sigmas = self.get_noise_schedule(n_steps, noise_scheduler or self.noise_scheduler)
        x = torch.randn((len(input_state), 1, self._core().config.act_dim), device=self.device) * self.sigma_max
        if len(self.action_context) > 0:                     # previously executed actions are re-denoised
            x = torch.cat([torch.cat(tuple(self.action_context), dim=1), x], dim=1)
        core = self._core()
        if io is not None:
            io.consumed, io.unscaled = False, None
            core.io_scaling = io
        try:
            x_0 = self.sample_loop(sigmas, x, input_state, goal, sampler_type, extra_args)
        finally:
            core.__dict__.pop("io_scaling", None)
        last = lambda v: v[:, -1, :] if (v.dim() == 3 and v.size(1) > 1) else v  # noqa: E731  (beso_agent.py:373-374)
        if io is not None and io.consumed:                  # clipped in the kernel; unscaled copy written beside it
            x_0, model_pred = last(x_0), last(io.unscaled)
        elif io is not None:                                 # step-wise loop (callback / churn): finish in torch
            x_0, model_pred = io.finish(last(x_0))
        else:
            if x_0.dim() == 3 and x_0.size(1) > 1:
                x_0 = x_0[:, -1, :]
            x_0 = self.scaler.clip_action(x_0)
            model_pred = self.scaler.inverse_scale_output(x_0)
        self._use_ema_weights(False)
        if model_pred.dim() == 2:
            x_0 = x_0.unsqueeze(1)
        self.action_context.append(x_0)
        return model_pred
