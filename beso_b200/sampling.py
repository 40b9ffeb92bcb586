"""Noise schedules and the DDIM / Euler / Heun samplers of the BESO sample loop.

Same call signatures as the reference's ``gc_sampling.sample_*`` (k_diffusion/gc_sampling.py) so
``BesoAgent.sample_loop`` (beso_agent.py:390-456) can dispatch to them unchanged.  When the model
is a beso_b200 ``GCDenoiser`` (optionally inside ``ClassifierFreeSampleModel``) and nothing needs
per-step Python (no callback, no scaler, no churn) the whole loop is ONE persistent kernel launch;
otherwise the loop below runs step by step and every ``model(...)`` call is one fused launch.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import _lib
from .denoiser import GCDenoiser


# ---- schedules (gc_sampling.py:22-95): O(n) host scalars, evaluated with the same fp32 torch ops
def append_zero(x: torch.Tensor) -> torch.Tensor:
    return torch.cat([x, x.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    return append_zero((hi + ramp * (lo - hi)) ** rho).to(device)


def get_sigmas_exponential(n, sigma_min, sigma_max, device="cpu"):
    return append_zero(torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp())


def get_sigmas_linear(n, sigma_min, sigma_max, device="cpu"):
    return append_zero(torch.linspace(sigma_max, sigma_min, n, device=device))


def get_sigmas_ve(n, sigma_min=0.02, sigma_max=100, device="cpu"):
    t = torch.linspace(0, n + 1, n, device=device)
    return append_zero(torch.sqrt((sigma_max ** 2) * ((sigma_min ** 2 / sigma_max ** 2) ** (t / (n - 1)))))


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, device="cpu"):
    t = torch.linspace(1, eps_s, n, device=device)
    return append_zero(torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1))


def get_sigmas_polyexponential(n, sigma_min, sigma_max, rho=1.0, device="cpu"):
    ramp = torch.linspace(1, 0, n, device=device) ** rho
    return append_zero(torch.exp(ramp * (math.log(sigma_max) - math.log(sigma_min)) + math.log(sigma_min)))


def to_d(action, sigma, denoised):
    """Karras ODE derivative (gc_sampling.py:98-100)."""
    return (action - denoised) / sigma.reshape(sigma.shape + (1,) * (action.ndim - sigma.ndim))


# ---- fused dispatch -----------------------------------------------------------------------
def _fusable(model):
    """Returns (denoiser, cfg_lambda, uncond) if ``model`` can run inside the persistent kernel."""
    from .cfg import ClassifierFreeSampleModel
    if isinstance(model, GCDenoiser):
        return model, None, False
    if isinstance(model, ClassifierFreeSampleModel) and isinstance(model.model, GCDenoiser):
        if model.cond:
            return model.model, None, False
        if model.cond_lambda == 0:
            return model.model, None, True
        return model.model, float(model.cond_lambda), False
    return None


def _try_fused(name, model, state, action, goal, sigmas, scaler, extra_args, callback, churn, coef=None, noise=None):
    if callback is not None or scaler is not None or churn or extra_args:
        return None
    f = _fusable(model)
    if f is None or not action.is_cuda or len(sigmas) < 2 or len(sigmas) > 129:
        return None
    den, lam, uncond = f
    if den.inner_model.training and den.inner_model.cond_mask_prob > 0:
        return None                       # goal masking draws from the RNG every call
    return den.sample(name, sigmas, state, action, goal, cfg_lambda=lam, uncond=uncond, coef=coef, noise=noise)


def ddim_coefficients(sigmas: torch.Tensor) -> torch.Tensor:
    """Per-step (sigma_fn(t_next)/sigma_fn(t), expm1(-h)) evaluated with the reference's own fp32
    tensor ops (gc_sampling.py:911-923) so the kernel applies bit-identical scalars."""
    s = sigmas.detach().float().cpu()
    t, t_next = s[:-1].log().neg(), s[1:].log().neg()
    h = t_next - t
    return torch.stack([t_next.neg().exp() / t.neg().exp(), (-h).expm1()], dim=1)


@torch.no_grad()
def sample_ddim(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                disable=None, eta=1.0):
    """DPM-Solver-1 / DDIM (gc_sampling.py:895-924); ``eta`` and ``scaler`` are ignored there too."""
    fused = _try_fused("ddim", model, state, action, goal, sigmas, None, extra_args, callback, 0.0,
                       coef=ddim_coefficients(sigmas))
    if fused is not None:
        return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"action": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        t, t_next = sigmas[i].log().neg(), sigmas[i + 1].log().neg()
        h = t_next - t
        action = (t_next.neg().exp() / t.neg().exp()) * action - (-h).expm1() * denoised
    return action


def _gamma(s_churn, n, sigma, s_tmin, s_tmax):
    return min(s_churn / n, 2 ** 0.5 - 1) if s_tmin <= sigma <= s_tmax else 0.0


@torch.no_grad()
def sample_euler(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                 disable=None, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, rng_parity=False):
    """Algorithm 2 of Karras et al. without the 2nd-order correction (gc_sampling.py:167-213).
    ``rng_parity`` keeps the reference's one ``randn_like`` draw per step in the fused path so the
    global RNG stream advances identically (results do not depend on it when s_churn == 0)."""
    fused = _try_fused("euler", model, state, action, goal, sigmas, scaler, extra_args, callback, s_churn)
    if fused is not None:
        if rng_parity:
            for _ in range(len(sigmas) - 1):
                torch.randn_like(action)
        return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    n = len(sigmas) - 1
    for i in range(n):
        gamma = _gamma(s_churn, n, sigmas[i], s_tmin, s_tmax)
        eps = torch.randn_like(action) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            action = action + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(state, action, goal, sigma_hat * ones, **extra_args)
        d = to_d(action, sigma_hat, denoised)
        if callback is not None:
            callback({"x": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "denoised": denoised})
        action = action + d * (sigmas[i + 1] - sigma_hat)
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


@torch.no_grad()
def sample_heun(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                disable=None, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, rng_parity=False):
    """Algorithm 2 of Karras et al. with Heun's correction; the step onto sigma = 0 is a plain Euler
    step (gc_sampling.py:259-314)."""
    fused = _try_fused("heun", model, state, action, goal, sigmas, scaler, extra_args, callback, s_churn)
    if fused is not None:
        if rng_parity:
            for _ in range(len(sigmas) - 1):
                torch.randn_like(action)
        return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    n = len(sigmas) - 1
    for i in range(n):
        gamma = _gamma(s_churn, n, sigmas[i], s_tmin, s_tmax)
        eps = torch.randn_like(action) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            action = action + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(state, action, goal, sigma_hat * ones, **extra_args)
        d = to_d(action, sigma_hat, denoised)
        if callback is not None:
            callback({"x": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "denoised": denoised})
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            action = action + d * dt
        else:
            action_2 = action + d * dt
            denoised_2 = model(state, action_2, goal, sigmas[i + 1] * ones, **extra_args)
            d_2 = to_d(action_2, sigmas[i + 1], denoised_2)
            action = action + ((d + d_2) / 2) * dt
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    """(sigma_down, sigma_up) of an ancestral step (gc_sampling.py:108-114), same expression on whatever the
    arguments are (0-d fp32 tensors in the samplers)."""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def ancestral_coefficients(sigmas: torch.Tensor, eta: float = 1.0) -> torch.Tensor:
    """Per-step (sigma_down, sigma_up) evaluated with the reference's own fp32 tensor ops on the host."""
    s = sigmas.detach().float().cpu()
    rows = []
    for i in range(len(s) - 1):
        down, up = get_ancestral_step(s[i], s[i + 1], eta=eta)
        rows.append([float(down), float(up)])
    return torch.tensor(rows, dtype=torch.float32)


@torch.no_grad()
def sample_euler_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                           disable=None, eta=1.0, noise=None):
    """Ancestral sampling with Euler steps (gc_sampling.py:216-256; the kitchen evaluation default,
    configs/evaluate_kitchen.yaml:12).  The ``torch.randn_like(action)`` draws of the reference -- one per step
    with ``sigma_down > 0``, in step order -- are made here from the same generator and handed to the persistent
    kernel, so the fused loop consumes the RNG stream exactly like the reference.  ``noise`` (n_steps, B, t, act)
    overrides the draws (parity tests)."""
    coef = ancestral_coefficients(sigmas, eta)
    n = len(sigmas) - 1
    fusable = (callback is None and scaler is None and not extra_args and _fusable(model) is not None and action.is_cuda
               and 2 <= len(sigmas) <= 129)
    if fusable:
        if noise is None:
            noise = torch.zeros((n,) + tuple(action.shape), device=action.device, dtype=torch.float32)
            for i in range(n):
                if float(coef[i, 0]) > 0:
                    noise[i] = torch.randn_like(action)
        fused = _try_fused("euler_ancestral", model, state, action, goal, sigmas, None, extra_args, callback, 0.0,
                           coef=coef, noise=noise)
        if fused is not None:
            return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    for i in range(n):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if callback is not None:
            callback({"x": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        d = to_d(action, sigmas[i], denoised)
        action = action + d * (sigma_down - sigmas[i])
        if sigma_down > 0:
            action = action + (noise[i] if noise is not None else torch.randn_like(action)) * sigma_up
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


def dpmpp_2m_coefficients(sigmas: torch.Tensor) -> torch.Tensor:
    """Per-step (sigma_fn(t_next)/sigma_fn(t), expm1(-h), 1 + 1/(2r), 1/(2r)) of DPM-Solver++(2M), evaluated with the
    reference's own fp32 tensor ops (gc_sampling.py:717-734); the last two are 0 on first-order steps."""
    s = sigmas.detach().float().cpu()
    t_fn = lambda sigma: sigma.log().neg()          # noqa: E731
    rows = []
    for i in range(len(s) - 1):
        t, t_next = t_fn(s[i]), t_fn(s[i + 1])
        h = t_next - t
        ca, ce = t_next.neg().exp() / t.neg().exp(), (-h).expm1()
        if i == 0 or s[i + 1] == 0:
            rows.append([float(ca), float(ce), 0.0, 0.0])
        else:
            h_last = t - t_fn(s[i - 1])
            r = h_last / h
            rows.append([float(ca), float(ce), float(1 + 1 / (2 * r)), float(1 / (2 * r))])
    return torch.tensor(rows, dtype=torch.float32)


@torch.no_grad()
def sample_dpmpp_2m(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None):
    """DPM-Solver++(2M) (gc_sampling.py:703-736); ``scaler`` is ignored there too."""
    fused = _try_fused("dpmpp_2m", model, state, action, goal, sigmas, None, extra_args, callback, 0.0,
                       coef=dpmpp_2m_coefficients(sigmas))
    if fused is not None:
        return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    t_fn = lambda sigma: sigma.log().neg()          # noqa: E731
    old_denoised = None
    for i in range(len(sigmas) - 1):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        if callback is not None:
            callback({"action": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        t, t_next = t_fn(sigmas[i]), t_fn(sigmas[i + 1])
        h = t_next - t
        if old_denoised is None or sigmas[i + 1] == 0:
            action = (t_next.neg().exp() / t.neg().exp()) * action - (-h).expm1() * denoised
        else:
            h_last = t - t_fn(sigmas[i - 1])
            r = h_last / h
            denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised
            action = (t_next.neg().exp() / t.neg().exp()) * action - (-h).expm1() * denoised_d
        old_denoised = denoised
    return action


# ---- second-order single-step samplers as a two-stage coefficient program (BESO_SAMPLER_TWO_STAGE) ---------------
# Every step is  D1 = model(x, sigma_i);  single-stage: x = a1 x + b1 D1 + su n_i;  two-stage: u = a1 x + b1 D1,
# D2 = model(u, sigma_b), x = a2 x + b2 u + c2 D2 + su n_i.  The rows below are [sigma_b, a1, b1, a2, b2, c2, su, 0],
# derived in float64 from the fp32 noise levels with the reference's formulas.
def _euler_row(sigma, target, su=0.0):
    dt = target - sigma                                    # x + (x - D)/sigma * dt
    return [0.0, 1.0 + dt / sigma, -dt / sigma, 0.0, 0.0, 0.0, su, 0.0]


def _dpm2_row(sigma, target, su=0.0):
    """DPM-Solver-2 step from sigma to target (gc_sampling.py:362-370, 400-408)."""
    mid = math.exp(0.5 * (math.log(sigma) + math.log(target)))          # log-space lerp at 0.5
    dt1, dt2 = mid - sigma, target - sigma
    return [mid, 1.0 + dt1 / sigma, -dt1 / sigma, 1.0, dt2 / mid, -dt2 / mid, su, 0.0]


def _dpmpp2s_row(sigma, target, su=0.0):
    """DPM-Solver++(2S) step from sigma to target (gc_sampling.py:958-964, 1004-1011), r = 1/2."""
    t, t_next = -math.log(sigma), -math.log(target)
    h = t_next - t
    s = t + 0.5 * h
    return [math.exp(-s), math.exp(-s) / math.exp(-t), -math.expm1(-h * 0.5), math.exp(-t_next) / math.exp(-t), 0.0,
            -math.expm1(-h), su, 0.0]


def two_stage_coefficients(kind: str, sigmas: torch.Tensor, eta: float = 1.0) -> torch.Tensor:
    s = [float(v) for v in sigmas.detach().float().cpu().tolist()]
    rows = []
    for i in range(len(s) - 1):
        if kind in ("dpm_2", "dpmpp_2s"):
            step = _dpm2_row if kind == "dpm_2" else _dpmpp2s_row
            rows.append(_euler_row(s[i], s[i + 1]) if s[i + 1] == 0 else step(s[i], s[i + 1]))
        elif kind in ("dpm_2_ancestral", "dpmpp_2s_ancestral"):
            down, up = get_ancestral_step(s[i], s[i + 1], eta=eta)
            step = _dpm2_row if kind == "dpm_2_ancestral" else _dpmpp2s_row
            rows.append(_euler_row(s[i], down, 0.0) if down == 0 else step(s[i], down, up))
        else:
            raise ValueError(f"unknown two-stage sampler {kind!r}")
    return torch.tensor(rows, dtype=torch.float32)


def _two_stage(kind, model, state, action, goal, sigmas, extra_args, callback, eta, noise, draws):
    """Fused path of the four samplers; ``draws`` = steps for which the reference calls randn_like (in order)."""
    coef = two_stage_coefficients(kind, sigmas, eta)
    n = len(sigmas) - 1
    if callback is not None or extra_args or _fusable(model) is None or not action.is_cuda or not (2 <= len(sigmas) <= 129):
        return None
    if noise is None and any(draws):
        noise = torch.zeros((n,) + tuple(action.shape), device=action.device, dtype=torch.float32)
        for i in range(n):
            if draws[i]:
                noise[i] = torch.randn_like(action)
    return _try_fused("two_stage", model, state, action, goal, sigmas, None, extra_args, callback, 0.0, coef=coef, noise=noise)


@torch.no_grad()
def sample_dpm_2(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                 s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, rng_parity=False):
    """DPM-Solver-2 (gc_sampling.py:317-377), s_churn = 0 in the fused path; the last step is an Euler step."""
    n = len(sigmas) - 1
    if not s_churn and scaler is None:
        fused = _two_stage("dpm_2", model, state, action, goal, sigmas, extra_args, callback, 1.0, None, [False] * n)
        if fused is not None:
            if rng_parity:                                  # the reference draws eps every step even when unused
                for _ in range(n):
                    torch.randn_like(action)
            return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    for i in range(n):
        gamma = _gamma(s_churn, n, sigmas[i], s_tmin, s_tmax)
        eps = torch.randn_like(action) * s_noise
        sigma_hat = sigmas[i] * (gamma + 1)
        if gamma > 0:
            action = action + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = model(state, action, goal, sigma_hat * ones, **extra_args)
        d = to_d(action, sigma_hat, denoised)
        if callback is not None:
            callback({"action": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigma_hat, "denoised": denoised})
        if sigmas[i + 1] == 0:
            action = action + d * (sigmas[i + 1] - sigma_hat)
        else:
            sigma_mid = sigma_hat.log().lerp(sigmas[i + 1].log(), 0.5).exp()
            action_2 = action + d * (sigma_mid - sigma_hat)
            denoised_2 = model(state, action_2, goal, sigma_mid * ones, **extra_args)
            action = action + to_d(action_2, sigma_mid, denoised_2) * (sigmas[i + 1] - sigma_hat)
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


@torch.no_grad()
def sample_dpm_2_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                           eta=1.0, noise=None):
    """Ancestral sampling with DPM-Solver-2 steps (gc_sampling.py:380-413): noise is drawn on the two-stage steps."""
    n = len(sigmas) - 1
    downs = [float(get_ancestral_step(float(sigmas[i]), float(sigmas[i + 1]), eta=eta)[0]) for i in range(n)]
    if scaler is None:
        fused = _two_stage("dpm_2_ancestral", model, state, action, goal, sigmas, extra_args, callback, eta, noise,
                           [d != 0 for d in downs])
        if fused is not None:
            return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    for i in range(n):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        if callback is not None:
            callback({"x": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        d = to_d(action, sigmas[i], denoised)
        if sigma_down == 0:
            action = action + d * (sigma_down - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigma_down.log(), 0.5).exp()
            action_2 = action + d * (sigma_mid - sigmas[i])
            denoised_2 = model(state, action_2, goal, sigma_mid * ones, **extra_args)
            action = action + to_d(action_2, sigma_mid, denoised_2) * (sigma_down - sigmas[i])
            action = action + (noise[i] if noise is not None else torch.randn_like(action)) * sigma_up
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


def _dpmpp_2s_loop(model, state, action, goal, sigmas, scaler, extra_args, callback, eta, s_noise, noise, ancestral):
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    sigma_fn = lambda t: t.neg().exp()              # noqa: E731
    t_fn = lambda sigma: sigma.log().neg()          # noqa: E731
    for i in range(len(sigmas) - 1):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        target, sigma_up = (get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta) if ancestral else (sigmas[i + 1], 0.0))
        if callback is not None:
            callback({"action": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        if target == 0:
            action = action + to_d(action, sigmas[i], denoised) * (target - sigmas[i])
        else:
            t, t_next = t_fn(sigmas[i]), t_fn(target)
            h = t_next - t
            s = t + 0.5 * h
            x_2 = (sigma_fn(s) / sigma_fn(t)) * action - (-h * 0.5).expm1() * denoised
            denoised_2 = model(state, x_2, goal, sigma_fn(s) * ones, **extra_args)
            action = (sigma_fn(t_next) / sigma_fn(t)) * action - (-h).expm1() * denoised_2
        if ancestral:                                       # the reference draws noise on EVERY step (sigma_up = 0 on the last)
            action = action + (noise[i] if noise is not None else torch.randn_like(action)) * s_noise * sigma_up
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


@torch.no_grad()
def sample_dpmpp_2s(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, eta=1.0):
    """DPM-Solver++(2S) (gc_sampling.py:928-967)."""
    if scaler is None:
        fused = _two_stage("dpmpp_2s", model, state, action, goal, sigmas, extra_args, callback, eta, None,
                           [False] * (len(sigmas) - 1))
        if fused is not None:
            return fused
    return _dpmpp_2s_loop(model, state, action, goal, sigmas, scaler, extra_args, callback, eta, 1.0, None, False)


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                              eta=1.0, s_noise=1.0, noise_sampler=None, noise=None):
    """Ancestral sampling with DPM-Solver++(2S) steps (gc_sampling.py:970-1016); default noise sampler only."""
    if scaler is None and noise_sampler is None and s_noise == 1.0:
        fused = _two_stage("dpmpp_2s_ancestral", model, state, action, goal, sigmas, extra_args, callback, eta, noise,
                           [True] * (len(sigmas) - 1))
        if fused is not None:
            return fused
    if noise_sampler is not None:
        raise NotImplementedError("custom noise samplers are not supported")
    return _dpmpp_2s_loop(model, state, action, goal, sigmas, scaler, extra_args, callback, eta, s_noise, noise, True)


def linear_multistep_coeff(order, t, i, j):
    """Coefficient of the j-th derivative of the i-th step of a linear multistep method (gc_sampling.py:416-428):
    the same scipy quadrature, with the same tolerance, as the reference."""
    from scipy import integrate
    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")

    def fn(tau):
        prod = 1.0
        for k in range(order):
            if j == k:
                continue
            prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod
    return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]


def lms_coefficients(sigmas: torch.Tensor, order: int = 4) -> torch.Tensor:
    """Per-step [c0, c1, c2, c3] of sample_lms; 0 for derivatives that do not exist yet."""
    if order > 4:
        raise ValueError("the fused linear multistep sampler supports order <= 4")
    t = sigmas.detach().cpu().numpy()
    rows = []
    for i in range(len(t) - 1):
        cur = min(i + 1, order)
        rows.append([linear_multistep_coeff(cur, t, i, j) for j in range(cur)] + [0.0] * (4 - cur))
    return torch.tensor(rows, dtype=torch.float32)


@torch.no_grad()
def sample_lms(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, order=4):
    """Linear multistep sampler (gc_sampling.py:431-468)."""
    f = _fusable(model)
    cfg_mix = f is not None and f[1] is not None
    if scaler is None and order <= 4 and not (cfg_mix and f[0].resolved_mode() == _lib.MODE_FAST):
        fused = _try_fused("lms", model, state, action, goal, sigmas, None, extra_args, callback, 0.0,
                           coef=lms_coefficients(sigmas, order))
        if fused is not None:
            return fused
    extra_args = {} if extra_args is None else extra_args
    ones = action.new_ones([action.shape[0]])
    sigmas_cpu = sigmas.detach().cpu().numpy()
    ds = []
    for i in range(len(sigmas) - 1):
        denoised = model(state, action, goal, sigmas[i] * ones, **extra_args)
        ds.append(to_d(action, sigmas[i], denoised))
        if len(ds) > order:
            ds.pop(0)
        if callback is not None:
            callback({"x": action, "i": i, "sigma": sigmas[i], "sigma_hat": sigmas[i], "denoised": denoised})
        cur_order = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur_order, sigmas_cpu, i, j) for j in range(cur_order)]
        action = action + sum(coeff * d for coeff, d in zip(coeffs, reversed(ds)))
        if scaler is not None:
            action = scaler.clip_output(action)
    return action


# ---- DPM-Solver-Fast / -Adaptive (gc_sampling.py:675-699, 855-892) ---------------------------------------------------
# Their step sequence is decided on the host (the adaptive one from an error norm of every trial step), so the loop
# cannot live in the persistent kernel; beso_b200/exp_integrator.py runs them as exponential Runge-Kutta tableaux with
# ONE fused denoiser launch per stage.  These two wrappers only carry the reference's call signatures.
@torch.no_grad()
def sample_dpm_fast(model, state, action, goal, sigma_min, sigma_max, n, scaler=None, extra_args=None, callback=None,
                    disable=None, eta=0.0, s_noise=1.0, noise_sampler=None):
    """DPM-Solver-Fast (gc_sampling.py:675-699); ``scaler`` and ``noise_sampler`` are ignored there too."""
    from .exp_integrator import integrate_fixed
    if sigma_min <= 0 or sigma_max <= 0:
        raise ValueError('sigma_min and sigma_max must not be 0')
    return integrate_fixed(model, state, action, goal, sigma_max, sigma_min, n, eta, s_noise, extra_args, callback)


@torch.no_grad()
def sample_dpm_adaptive(model, state, action, goal, sigma_min, sigma_max, extra_args=None, callback=None, disable=None,
                        order=3, rtol=0.05, atol=0.0078, h_init=0.05, pcoeff=0.0, icoeff=1.0, dcoeff=0.0,
                        accept_safety=0.81, eta=0.0, s_noise=1.0, return_info=False):
    """DPM-Solver-12 / 23 with adaptive step size (gc_sampling.py:855-892)."""
    from .exp_integrator import integrate_adaptive
    if sigma_min <= 0 or sigma_max <= 0:
        raise ValueError('sigma_min and sigma_max must not be 0')
    action, info = integrate_adaptive(model, state, action, goal, sigma_max, sigma_min, order, rtol, atol, h_init,
                                      (pcoeff, icoeff, dcoeff), accept_safety, eta, s_noise, extra_args, callback)
    return (action, info) if return_info else action


SAMPLERS = {"ddim": sample_ddim, "euler": sample_euler, "heun": sample_heun, "euler_ancestral": sample_euler_ancestral,
            "lms": sample_lms, "dpmpp_2m": sample_dpmpp_2m, "dpm": sample_dpm_2, "ancestral": sample_dpm_2_ancestral,
            "dpmpp_2s": sample_dpmpp_2s, "dpmpp_2s_ancestral": sample_dpmpp_2s_ancestral}


def n_model_evals(sampler: str, sigmas) -> int:
    """Model evaluations per sequence for one loop (second-order single-step samplers: 2 per step, 1 on a final
    sigma = 0 step; the ancestral variants also take a single evaluation when sigma_down = 0)."""
    n = len(sigmas) - 1
    if sampler in ("heun", "dpm", "dpmpp_2s"):
        return 2 * n - (1 if float(sigmas[-1]) == 0.0 else 0)
    if sampler in ("ancestral", "dpmpp_2s_ancestral"):
        return n + sum(1 for i in range(n) if float(get_ancestral_step(float(sigmas[i]), float(sigmas[i + 1]))[0]) != 0.0)
    return n
