"""DPM-Solver fast / adaptive (gc_sampling.py:498-699, 855-892): host-side step control over a model callable.
CPU: the exponential Runge-Kutta tableaux of beso_b200/exp_integrator.py driven by the oracle forward reproduce the
reference's outputs (fixtures made with the real reference, oracle/make_golden.py dpm_solver; live comparison when
/root/reference is present) -- to fp32 round-off, the tableau sums its terms in a different order than the reference.
GPU: the same loops over the fused denoiser (every evaluation one launch) through BesoAgent.sample_loop."""
import os

import pytest
import torch

from conftest import golden_weights, load_golden, to_oracle_cfg
from beso_b200 import sampling as S
from oracle import beso_oracle as O
from oracle import ref_import

TOL = dict(rtol=2e-5, atol=2e-6)


def _oracle_model(cfg, meta):
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    return lambda state, action, goal, sigma, **kw: O.denoiser_forward(sd, oc, state, action, goal, sigma, **kw)


def test_dpm_fast_matches_reference_golden():
    cfg, meta, a = load_golden("samplers_dpm_solver_K256")
    model = _oracle_model(cfg, meta)
    for n in (3, 6, 8, 10):
        torch.manual_seed(7500 + n)
        got = S.sample_dpm_fast(model, a["state"], a["x_t"], a["goal"], 0.005, 1.0, n, disable=True)
        torch.testing.assert_close(got, a[f"dpm_fast_{n}"], **TOL)
    # eta > 0: ancestral steps, one randn_like per solver step in the reference's order
    torch.manual_seed(7600)
    got = S.sample_dpm_fast(model, a["state"], a["x_t"], a["goal"], 0.005, 1.0, 9, disable=True, eta=0.5)
    torch.testing.assert_close(got, a["dpm_fast_eta_9"], **TOL)
    with pytest.raises(ValueError):
        S.sample_dpm_fast(model, a["state"], a["x_t"], a["goal"], 0.0, 1.0, 3)


def test_dpm_adaptive_matches_reference_golden():
    cfg, meta, a = load_golden("samplers_dpm_solver_K256")
    model = _oracle_model(cfg, meta)
    for order in (2, 3):
        torch.manual_seed(7700 + order)
        got, info = S.sample_dpm_adaptive(model, a["state"], a["x_t"], a["goal"], 0.005, 1.0, disable=True, order=order,
                                          return_info=True)
        want = a[f"dpm_adaptive_{order}_info"]
        assert [info[k] for k in ("steps", "nfe", "n_accept", "n_reject")] == want.tolist()   # same accept / reject path
        torch.testing.assert_close(got, a[f"dpm_adaptive_{order}"], rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        S.sample_dpm_adaptive(model, a["state"], a["x_t"], a["goal"], 0.005, 1.0, order=4)


def test_step_size_controller_and_tableaux():
    from beso_b200.exp_integrator import _pid, tableau
    box, propose = _pid(0.05, (0.0, 1.0, 0.0), 3, 0.81)
    assert propose(0.5) and box["h"] > 0.05                  # small error: accepted, step grows
    h = box["h"]
    assert not propose(50.0) and box["h"] < h                # large error: rejected, step shrinks
    # consistency of the tableaux: for a noise prediction that does not depend on (u, s) every order reduces to the
    # first-order (exponential Euler) step, i.e. the weights of each order sum to the first-order weight
    t, tn = torch.tensor(0.3), torch.tensor(1.1)
    w1 = tableau(1, t, tn)[2][0]
    for order in (2, 3):
        nodes, A, b = tableau(order, t, tn)
        assert len(nodes) == order and all(len(A[k]) == k for k in range(order))
        torch.testing.assert_close(sum(b), w1, rtol=1e-6, atol=1e-7)
        for k in range(1, order):                            # ... and so does every stage: u_k is the first-order step to s_k
            torch.testing.assert_close(sum(A[k]), tableau(1, t, nodes[k])[2][0], rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_dpm_solver_matches_live_reference_with_callback():
    """Same model object through both implementations: equal to fp32 round-off (same formulas, different summation
    order), same accept / reject path of the adaptive solver, same callback payload keys."""
    from beso_b200.config import ModelConfig
    from beso_b200.synth import synthetic_inputs, synthetic_state_dict
    ns = ref_import.load()
    cfg = ModelConfig(obs_dim=12, act_dim=3, window=3, goal_len=1, d=32, n_layers=2, n_heads=2)
    m = ref_import.make_reference_model(ns, cfg)
    m.load_state_dict(synthetic_state_dict(cfg, 3))
    m.eval()
    x = synthetic_inputs(cfg, 5, seed=4)
    seen_a, seen_b = [], []
    for n in (4, 9):
        torch.manual_seed(n)
        want = ns.gc_sampling.sample_dpm_fast(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, n, disable=True,
                                              callback=lambda i: seen_a.append(sorted(i)))
        torch.manual_seed(n)
        got = S.sample_dpm_fast(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, n, disable=True,
                                callback=lambda i: seen_b.append(sorted(i)))
        torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-6)
    assert seen_a == seen_b and len(seen_a) > 0
    for kw in (dict(order=3), dict(order=2, rtol=0.02), dict(order=3, eta=0.3)):
        torch.manual_seed(11)
        want, wi = ns.gc_sampling.sample_dpm_adaptive(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, disable=True,
                                                      return_info=True, **kw)
        torch.manual_seed(11)
        got, gi = S.sample_dpm_adaptive(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, disable=True, return_info=True, **kw)
        assert gi == wi
        torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_agent_dispatches_dpm_solvers_gpu(cuda_device):
    """BesoAgent.sample_loop('dpm_fast' / 'dpm_adaptive') over the fused denoiser against the reference goldens."""
    from beso_b200.agent import BesoAgent
    from beso_b200.denoiser import build_denoiser
    cfg, meta, a = load_golden("samplers_dpm_solver_K256")
    m = build_denoiser(cfg, cuda_device, mode="precise", state_dict=golden_weights(cfg, meta))
    m.eval()
    agent = BesoAgent(m, device=cuda_device, sigma_min=0.005, sigma_max=1.0, window_size=cfg.window)
    dev = {k: a[k].to(cuda_device) for k in ("state", "goal", "x_t")}
    for n in (6, 10):                                        # the agent passes len(sigmas) as the evaluation budget
        sigmas = agent.get_noise_schedule(n - 1, "exponential")
        assert len(sigmas) == n and abs(float(sigmas[-2]) - 0.005) < 1e-7
        got = agent.sample_loop(sigmas, dev["x_t"], dev["state"], dev["goal"], "dpm_fast")
        torch.testing.assert_close(got.cpu(), a[f"dpm_fast_{n}"], rtol=1e-3, atol=2e-5)
    sigmas = agent.get_noise_schedule(5, "exponential")
    got = agent.sample_loop(sigmas, dev["x_t"], dev["state"], dev["goal"], "dpm_adaptive")
    torch.testing.assert_close(got.cpu(), a["dpm_adaptive_3"], rtol=1e-3, atol=5e-5)
