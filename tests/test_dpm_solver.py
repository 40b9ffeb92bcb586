"""DPM-Solver fast / adaptive (gc_sampling.py:498-699, 855-892): host-side step control over a model callable.
CPU: the loops of beso_b200.sampling driven by the oracle forward reproduce the reference's outputs (fixtures made with
the real reference, oracle/make_golden.py dpm_solver; live comparison when /root/reference is present).
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


def test_step_size_controller():
    pid = S.PIDStepSizeController(0.05, 0.0, 1.0, 0.0, order=3, accept_safety=0.81)
    assert pid.propose_step(0.5) and pid.h > 0.05            # small error: accepted, step grows
    h = pid.h
    assert not pid.propose_step(50.0) and pid.h < h          # large error: rejected, step shrinks
    assert pid.limiter(1.0) == 1.0


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_dpm_solver_matches_live_reference_with_callback():
    """Same model object through both implementations: bit-identical, including the callback payload keys."""
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
        assert torch.equal(got, want)
    assert seen_a == seen_b and len(seen_a) > 0
    for kw in (dict(order=3), dict(order=2, rtol=0.02), dict(order=3, eta=0.3)):
        torch.manual_seed(11)
        want, wi = ns.gc_sampling.sample_dpm_adaptive(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, disable=True,
                                                      return_info=True, **kw)
        torch.manual_seed(11)
        got, gi = S.sample_dpm_adaptive(m, x["state"], x["noise"], x["goal"], 0.01, 1.0, disable=True, return_info=True, **kw)
        assert gi == wi and torch.equal(got, want)


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("BESO_RUN_UNVERIFIED_GPU_TESTS") != "1",
                    reason="written after this round's GPU minutes were spent: never run on a GPU box yet; "
                           "set BESO_RUN_UNVERIFIED_GPU_TESTS=1 to run it")
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
