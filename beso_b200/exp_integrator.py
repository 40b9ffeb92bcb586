"""Host-driven exponential integrators for the two samplers whose step sequence is data dependent or irregular:
``dpm_fast`` and ``dpm_adaptive`` (reference behaviour: k_diffusion/gc_sampling.py:527-699, 855-892; dispatched by
BesoAgent.sample_loop, beso_agent.py:438-442).  They are outside SURVEY.md section 8's fused scope -- the step sizes of
the adaptive solver depend on an error norm of every trial step -- so the loop runs on the host and every model
evaluation is ONE fused launch of the denoiser.

Formulation used here.  In log-noise time t = -log(sigma) the probability-flow ODE of the denoiser D is
``dx/dt = -(x - D(x, t))``; with the noise prediction ``e(u, s) = (u - D(u, sigma(s))) / sigma(s)`` a DPM-Solver step of
order p from t to t + h is an explicit exponential Runge-Kutta scheme: stage k evaluates ``e_k = e(u_k, s_k)`` at
``u_k = x + sum_{j<k} A[k][j] e_j`` and the step ends at ``x + sum_j b[j] e_j``.  ``tableau()`` writes the published
DPM-Solver-1/2/3 updates (Lu et al. 2022) in that form once; the fixed-step solver and the embedded-pair adaptive solver
are then the same few lines over a tableau: stages shared between the members of a pair are evaluated once, the PID
step-size controller only sees scalars.
"""
from __future__ import annotations

import math

import torch

from .sampling import get_ancestral_step


def _sigma(t):
    return t.neg().exp()


def tableau(order: int, t, t_next, r1=None, r2=None):
    """Nodes ``s`` (stage times), strictly lower-triangular ``A`` and weights ``b`` of the DPM-Solver step of the given
    order from t to t_next (0-d tensors), all as 0-d tensors so that the arithmetic stays in fp32 on the host.

    order 1:  x' = x - sigma' expm1(h) e0
    order 2:  u1 = x - sigma(s1) expm1(r1 h) e0,   x' = x - sigma' expm1(h) e0 - sigma' expm1(h) / (2 r1) (e1 - e0)
    order 3:  u1 as above, u2 = x - sigma(s2) expm1(r2 h) e0 - sigma(s2) (r2 / r1) (expm1(r2 h) / (r2 h) - 1) (e1 - e0),
              x' = x - sigma' expm1(h) e0 - sigma' / r2 (expm1(h) / h - 1) (e2 - e0)
    """
    h = t_next - t
    sn = _sigma(t_next)
    zero = torch.zeros_like(h)
    lead = sn * h.expm1()                                    # weight of e0 in every order
    if order == 1:
        return [t], [[]], [-lead]
    if order == 2:
        r1 = 0.5 if r1 is None else r1
        s1 = t + r1 * h
        c = lead / (2 * r1)
        return [t, s1], [[], [-_sigma(s1) * (r1 * h).expm1()]], [c - lead, -c]
    r1 = 1 / 3 if r1 is None else r1
    r2 = 2 / 3 if r2 is None else r2
    s1, s2 = t + r1 * h, t + r2 * h
    g2 = _sigma(s2) * (r2 / r1) * ((r2 * h).expm1() / (r2 * h) - 1)
    c = sn / r2 * (h.expm1() / h - 1)
    return ([t, s1, s2],
            [[], [-_sigma(s1) * (r1 * h).expm1()], [g2 - _sigma(s2) * (r2 * h).expm1(), -g2]],
            [c - lead, zero, -c])


class _Stages:
    """Evaluates the noise predictions of one step lazily and once: members of an embedded pair share their leading
    stages (same node, same ``u``)."""

    def __init__(self, model, state, goal, extra_args):
        self.model, self.state, self.goal = model, state, goal
        self.extra = {} if extra_args is None else extra_args
        self.cache = {}
        self.n_evals = 0

    def noise_pred(self, key, u, s):
        if key not in self.cache:
            sig = _sigma(s)
            d = self.model(self.state, u, self.goal, sig * u.new_ones([u.shape[0]]), **self.extra)
            self.cache[key] = (u - d) / sig
            self.n_evals += 1
        return self.cache[key]

    def run(self, x, nodes, A, b, tag=""):
        """x' of one tableau; stage k is cached under (k, tag-of-its-row) -- rows that coincide between two tableaux
        are given the same tag by the caller."""
        e = []
        for k, s in enumerate(nodes):
            u = x
            for j, a in enumerate(A[k]):
                u = u + a * e[j]
            e.append(self.noise_pred((k, tag if k > 1 else ""), u, s))
        out = x
        for w, ek in zip(b, e):
            out = out + w * ek
        return out


def _ancestral_target(t, t_next, t_end, eta):
    """(time the deterministic part of the step goes to, sigma_up of the noise added afterwards); eta = 0: plain ODE."""
    if not eta:
        return t_next, 0.0
    sd, _ = get_ancestral_step(_sigma(t), _sigma(t_next), eta)
    t_to = torch.minimum(t_end, -sd.log())
    return t_to, (_sigma(t_next) ** 2 - _sigma(t_to) ** 2) ** 0.5


def _report(callback, x, i, t, stages, extra=None):
    if callback is None:
        return
    e0 = stages.cache[(0, "")]
    info = {'x': x, 'i': i, 't': t, 't_up': t, 'denoised': x - _sigma(t) * e0, 'sigma': _sigma(t), 'sigma_hat': _sigma(t)}
    info.update(extra or {})
    callback(info)


def integrate_fixed(model, state, x, goal, sigma_start, sigma_end, nfe, eta=0.0, s_noise=1.0, extra_args=None, callback=None):
    """``nfe`` model evaluations over uniform steps in t: third-order steps with a second / first order remainder."""
    t0, t1 = -torch.tensor(float(sigma_start)).log(), -torch.tensor(float(sigma_end)).log()
    if not t1 > t0 and eta:
        raise ValueError('eta must be 0 for reverse sampling')
    m = nfe // 3 + 1
    grid = torch.linspace(t0, t1, m + 1, device=x.device)
    orders = [3] * (m - 2) + [2, 1] if nfe % 3 == 0 else [3] * (m - 1) + [nfe % 3]
    for i, order in enumerate(orders):
        t, t_next = grid[i], grid[i + 1]
        t_to, s_up = _ancestral_target(t, t_next, t1, eta)
        st = _Stages(model, state, goal, extra_args)
        nodes, A, b = tableau(order, t, t_to)
        st.noise_pred((0, ""), x, t)
        _report(callback, x, i, t, st)
        x = st.run(x, nodes, A, b)
        x = x + s_up * s_noise * torch.randn_like(x)          # the draw is made even when s_up = 0 (RNG parity)
    return x


def _pid(h, coeffs, order, accept_safety, eps=1e-8):
    """PID step-size controller (Soderlind): ``propose(error)`` scales h by 1 + atan(f - 1) with
    f = prod_k (1 / (error_k + eps)) ** beta_k over the last three errors and accepts the step if that factor is at
    least ``accept_safety``."""
    kp, ki, kd = coeffs
    beta = ((kp + ki + kd) / order, -(kp + 2 * kd) / order, kd / order)
    hist = []
    box = {"h": h}

    def propose(error):
        inv = 1 / (float(error) + eps)
        if not hist:
            hist.extend([inv, inv, inv])
        hist[0] = inv
        factor = 1 + math.atan(hist[0] ** beta[0] * hist[1] ** beta[1] * hist[2] ** beta[2] - 1)
        ok = factor >= accept_safety
        if ok:
            hist[2], hist[1] = hist[1], hist[0]
        box["h"] *= factor
        return ok
    return box, propose


def integrate_adaptive(model, state, x, goal, sigma_start, sigma_end, order=3, rtol=0.05, atol=0.0078, h_init=0.05,
                       pid_coeffs=(0.0, 1.0, 0.0), accept_safety=0.81, eta=0.0, s_noise=1.0, extra_args=None, callback=None):
    """Embedded pairs DPM-Solver-1(2) / 2(3): a trial step is taken with both members, the difference measured
    against ``max(atol, rtol max(|low|, |previous low|))`` and the step size driven by the PID controller."""
    if order not in {2, 3}:
        raise ValueError('order should be 2 or 3')
    t0, t1 = -torch.tensor(float(sigma_start)).log(), -torch.tensor(float(sigma_end)).log()
    forward = bool(t1 > t0)
    if not forward and eta:
        raise ValueError('eta must be 0 for reverse sampling')
    box, propose = _pid(abs(h_init) * (1 if forward else -1), pid_coeffs, 1.5 if eta else order, accept_safety)
    atol_t, rtol_t = torch.tensor(atol), torch.tensor(rtol)
    s, x_prev = t0, x
    info = {'steps': 0, 'nfe': 0, 'n_accept': 0, 'n_reject': 0}
    while (s < t1 - 1e-5) if forward else (s > t1 + 1e-5):
        t = torch.minimum(t1, s + box["h"]) if forward else torch.maximum(t1, s + box["h"])
        t_to, s_up = _ancestral_target(s, t, t1, eta)
        st = _Stages(model, state, goal, extra_args)
        st.noise_pred((0, ""), x, s)
        denoised = x - _sigma(s) * st.cache[(0, "")]
        if order == 2:
            low = st.run(x, *tableau(1, s, t_to))
            high = st.run(x, *tableau(2, s, t_to))
        else:                                                # the pair shares e0 and the r1 = 1/3 stage
            low = st.run(x, *tableau(2, s, t_to, r1=1 / 3))
            high = st.run(x, *tableau(3, s, t_to))
        delta = torch.maximum(atol_t.to(low.device), rtol_t.to(low.device) * torch.maximum(low.abs(), x_prev.abs()))
        error = torch.linalg.norm((low - high) / delta) / x.numel() ** 0.5
        if propose(error):
            x_prev = low
            x = high + s_up * s_noise * torch.randn_like(x)
            s = t
            info['n_accept'] += 1
        else:
            info['n_reject'] += 1
        info['nfe'] += order
        info['steps'] += 1
        if callback is not None:
            callback({'x': x, 'i': info['steps'] - 1, 't': s, 't_up': s, 'denoised': denoised, 'error': error, 'h': box["h"],
                      'sigma': _sigma(s), 'sigma_hat': _sigma(s), **info})
    return x, info
