"""CPU: the oracle restatement against (a) the committed golden vectors produced by the
real reference and (b) the live reference incl. all shipped checkpoints when it is present."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden, to_oracle_cfg
from oracle import beso_oracle as O
from oracle import ref_import

# The oracle replays the reference's ATen ops in order; on the torch build / thread count that
# made the fixtures it is bit-exact (asserted against the live reference below).  The tolerance
# only absorbs a different BLAS thread split on another host.
TOL = dict(rtol=1e-5, atol=1e-6)

FWD = ["fwd_K256", "fwd_K256_t1", "fwd_K256_t4", "fwd_T16", "fwd_B256", "fwd_small_kitchen",
       "fwd_small_push", "fwd_mlp_head"]


@pytest.mark.parametrize("name", FWD)
def test_forward_matches_golden(name):
    cfg, meta, a = load_golden(name)
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        out = O.denoiser_forward(sd, oc, a["state"], a["action"], a["goal"], a["sigma"])
        out_u = O.denoiser_forward(sd, oc, a["state"], a["action"], a["goal"], a["sigma"], uncond=True)
        c_in = O.get_scalings(a["sigma"], oc.sigma_data)[2]
    torch.testing.assert_close(out, a["out"], **TOL)
    torch.testing.assert_close(out_u, a["out_uncond"], **TOL)
    assert not torch.equal(out, out_u)
    del c_in


def test_forward_no_goal_conditioning():
    cfg, meta, a = load_golden("fwd_no_goal")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        out = O.denoiser_forward(sd, oc, a["state"], a["action"], a["goal"], a["sigma"])
    torch.testing.assert_close(out, a["out"], **TOL)


def test_samplers_match_golden():
    cfg, meta, a = load_golden("samplers_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        for n in (1, 3, 5):
            sig = a[f"sigmas_{n}"]
            for s in ("ddim", "euler", "heun"):
                got = O.SAMPLERS[s](sd, oc, a["state"], a["x_t"], a["goal"], sig)
                torch.testing.assert_close(got, a[f"{s}_{n}"], **TOL)
        got = O.sample_heun(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_karras_4"])
        torch.testing.assert_close(got, a["heun_karras_4"], **TOL)
        for lam in (0.0, 1.0, 1.5, 2.0):
            tag = str(lam).replace(".", "p")
            got = O.cfg_forward(sd, oc, lam, a["state"], a["action"], a["goal"], a["sigma"])
            torch.testing.assert_close(got, a[f"cfg_fwd_{tag}"], **TOL)
            got = O.sample_heun(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_cfg_4"], cond_lambda=lam)
            torch.testing.assert_close(got, a[f"cfg_heun4_{tag}"], **TOL)
            got = O.sample_ddim(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_cfg_4"], cond_lambda=lam)
            torch.testing.assert_close(got, a[f"cfg_ddim4_{tag}"], **TOL)


def test_euler_ancestral_matches_golden():
    """The reference's sample_euler_ancestral under a fixed generator state (oracle/make_golden.py ancestral):
    the oracle reproduces it both from the recorded draws and by re-seeding the generator."""
    cfg, meta, a = load_golden("samplers_ancestral_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        for n in (1, 3, 5):
            got = O.sample_euler_ancestral(sd, oc, a["state"], a["x_t"], a["goal"], a[f"sigmas_{n}"], noise=a[f"noise_{n}"])
            torch.testing.assert_close(got, a[f"euler_ancestral_{n}"], **TOL)
            torch.manual_seed(7000 + n)
            got = O.sample_euler_ancestral(sd, oc, a["state"], a["x_t"], a["goal"], a[f"sigmas_{n}"])
            torch.testing.assert_close(got, a[f"euler_ancestral_{n}"], **TOL)
        got = O.sample_euler_ancestral(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_karras_4"], noise=a["noise_karras_4"])
        torch.testing.assert_close(got, a["euler_ancestral_karras_4"], **TOL)
    # the step onto sigma = 0 draws no noise and lands on the denoised sample
    assert float(a["noise_3"][-1].abs().max()) == 0.0


def test_dpmpp_2m_matches_golden():
    """DPM-Solver++(2M) (gc_sampling.py:703-736) of the real reference."""
    cfg, meta, a = load_golden("samplers_ancestral_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        for tag in ("1", "3", "5", "karras_4"):
            got = O.sample_dpmpp_2m(sd, oc, a["state"], a["x_t"], a["goal"], a[f"sigmas_{tag}"])
            torch.testing.assert_close(got, a[f"dpmpp_2m_{tag}"], **TOL)


def test_ddim_last_step_returns_denoised():
    """sigma_{n}=0 -> h=+inf -> x <- denoised exactly (gc_sampling.py:921-923)."""
    cfg, meta, a = load_golden("samplers_K256")
    torch.testing.assert_close(a["ddim_5"], a["ddim_5_denoised_trace"][-1], rtol=0, atol=0)
    torch.testing.assert_close(a["ddim_1"], a["euler_1"], rtol=1e-5, atol=1e-6)   # one step to sigma=0
    torch.testing.assert_close(a["heun_1"], a["euler_1"], rtol=0, atol=0)         # Heun's last step is Euler


def test_schedules_match_golden():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "schedules.npz"))
    for n in (1, 3, 10, 50):
        np.testing.assert_array_equal(O.get_sigmas_exponential(n, 0.005, 1.0).numpy(), z[f"exponential_{n}"])
        np.testing.assert_array_equal(O.get_sigmas_karras(n, 0.005, 1.0, 5.0).numpy(), z[f"karras_{n}"])
        np.testing.assert_array_equal(O.get_sigmas_linear(n, 0.005, 1.0).numpy(), z[f"linear_{n}"])
        np.testing.assert_array_equal(O.get_sigmas_vp(n).numpy(), z[f"vp_{n}"])
        if n > 1:
            np.testing.assert_array_equal(O.get_sigmas_ve(n, 0.005, 1.0).numpy(), z[f"ve_{n}"])


@pytest.mark.parametrize("name", ["loss_B256", "loss_K256"])
def test_loss_and_grads_match_golden(name):
    cfg, meta, a = load_golden(name)
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    loss, grads = O.loss_and_grads(sd, oc, a["state"], a["action"], a["goal"], a["noise"].clone(), a["sigma"])
    torch.testing.assert_close(loss, a["loss"], rtol=1e-5, atol=1e-7)
    names = [str(n) for n in a["grad_names"]]
    for n, ref_norm in zip(names, a["grad_norms"]):
        g = grads["inner_model." + n if not n.startswith("inner_model.") else n]
        assert abs(g.double().norm().item() - ref_norm) <= 1e-4 * ref_norm + 1e-9, n
        flat = g.reshape(-1)
        want = a["grad::" + n]
        got = flat if flat.numel() <= 4096 else flat[::97][:4096]
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-7)
    with torch.no_grad():
        l2 = O.denoiser_loss(sd, oc, a["state"], a["action"], a["goal"], a["noise"].clone(), a["sigma"],
                             pred_last_action_only=True)
    torch.testing.assert_close(l2, a["loss_pred_last"], rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------
# live reference (build container only)
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")


@needs_ref
@pytest.mark.parametrize("env", ["kitchen", "block_push"])
def test_oracle_bitexact_on_shipped_checkpoints(env):
    from beso_b200.config import BLOCKPUSH_CKPT, KITCHEN_CKPT
    from beso_b200.synth import synthetic_inputs
    cfg = KITCHEN_CKPT if env == "kitchen" else BLOCKPUSH_CKPT
    ns = ref_import.load()
    paths = sorted(glob.glob(os.path.join(ref_import.REF_ROOT, "trained_models", env, "*", "*.pth")))
    assert len(paths) == 12
    oc = to_oracle_cfg(cfg)
    x = synthetic_inputs(cfg, 6, seed=5)
    sig = ns.gc_sampling.get_sigmas_exponential(3, 0.005, 1.0)
    for p in paths[::3]:
        sd = torch.load(p, map_location="cpu")
        m = ref_import.make_reference_model(ns, cfg)
        m.load_state_dict(sd, strict=True)
        sd = O.as_module_params(sd)
        with torch.no_grad():
            want = m(x["state"], x["action"], x["goal"], x["sigma"])
            got = O.denoiser_forward(sd, oc, x["state"], x["action"], x["goal"], x["sigma"])
            assert torch.equal(want, got), p
            want = ns.gc_sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig, disable=True)
            got = O.sample_ddim(sd, oc, x["state"], x["noise"], x["goal"], sig)
            assert torch.equal(want, got), p


def _run_two_stage_program(model, state, x, goal, sigmas, coef, noise):
    """CPU emulation of BESO_SAMPLER_TWO_STAGE (include/beso_b200.h) on top of any model(state, x, goal, sigma)."""
    ones = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        sb, a1, b1, a2, b2, c2, su, _ = [float(v) for v in coef[i]]
        d1 = model(state, x, goal, sigmas[i] * ones)
        nz = noise[i] * su if su != 0.0 else 0.0
        if sb == 0.0:
            x = a1 * x + b1 * d1 + nz
        else:
            u = a1 * x + b1 * d1
            d2 = model(state, u, goal, torch.tensor(sb) * ones)
            x = a2 * x + b2 * u + c2 * d2 + nz
    return x


def test_second_order_samplers_match_golden_and_their_coefficient_programs():
    """dpm_2 / dpm_2_ancestral / dpmpp_2s / dpmpp_2s_ancestral: the oracle restatements reproduce the real reference,
    and the two-stage coefficient programs the CUDA kernels execute (beso_b200.sampling.two_stage_coefficients)
    describe the same samplers."""
    from beso_b200 import sampling
    cfg, meta, a = load_golden("samplers_ancestral_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    model = lambda s, x, g, sig: O.denoiser_forward(sd, oc, s, x, g, sig)     # noqa: E731
    with torch.no_grad():
        for tag in ("3", "5", "karras_4"):
            sig = a[f"sigmas_{tag}"]
            zeros = torch.zeros((len(sig) - 1,) + tuple(a["x_t"].shape))
            cases = {
                "dpm_2": (O.sample_dpm_2(sd, oc, a["state"], a["x_t"], a["goal"], sig), zeros),
                "dpmpp_2s": (O.sample_dpmpp_2s(sd, oc, a["state"], a["x_t"], a["goal"], sig), zeros),
                "dpm_2_ancestral": (O.sample_dpm_2_ancestral(sd, oc, a["state"], a["x_t"], a["goal"], sig,
                                                             noise=a[f"noise_dpm2a_{tag}"]), a[f"noise_dpm2a_{tag}"]),
                "dpmpp_2s_ancestral": (O.sample_dpmpp_2s(sd, oc, a["state"], a["x_t"], a["goal"], sig, ancestral=True,
                                                         noise=a[f"noise_2sa_{tag}"]), a[f"noise_2sa_{tag}"]),
            }
            for kind, (got, noise) in cases.items():
                torch.testing.assert_close(got, a[f"{kind}_{tag}"], **TOL)
                prog = _run_two_stage_program(model, a["state"], a["x_t"], a["goal"], sig,
                                              sampling.two_stage_coefficients(kind, sig), noise)
                torch.testing.assert_close(prog, a[f"{kind}_{tag}"], rtol=1e-4, atol=2e-6)


def test_lms_matches_golden():
    """Linear multistep sampler (gc_sampling.py:416-468) of the real reference, orders 1..4."""
    from beso_b200 import sampling
    cfg, meta, a = load_golden("samplers_ancestral_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    with torch.no_grad():
        for tag in ("3", "6", "karras_4"):
            got = O.sample_lms(sd, oc, a["state"], a["x_t"], a["goal"], a[f"sigmas_{tag}"])
            torch.testing.assert_close(got, a[f"lms_{tag}"], **TOL)
    coef = sampling.lms_coefficients(a["sigmas_6"])
    assert coef.shape == (6, 4) and float(coef[0, 1]) == 0.0 and float(coef[2, 3]) == 0.0 and float(coef[3, 3]) != 0.0
    t = a["sigmas_6"].numpy()
    assert float(coef[4, 2]) == pytest.approx(O.linear_multistep_coeff(4, t, 4, 2), rel=1e-6)


# ---- training-mode dropout: the product's mask draws replay the reference's generator consumption ----------------
def _replay_training_draws(name):
    """Re-draws, under the fixture's seed on the CPU, what the reference drew inside GCDenoiser.loss: the CFG goal mask
    (score_gpts.py:366) and then the nn.Dropout masks in op order (beso_b200.training.draw_dropout_masks)."""
    from conftest import golden_weights, load_golden
    from beso_b200.denoiser import build_denoiser
    from beso_b200.training import draw_dropout_masks
    cfg, meta, a = load_golden(name)
    sd = golden_weights(cfg, meta)
    m = build_denoiser(cfg, "cpu", state_dict=sd, attn_pdrop=float(a["attn_pdrop"]), resid_pdrop=float(a["resid_pdrop"]),
                       goal_drop=float(a["goal_drop"]))
    m.train()
    torch.manual_seed(int(a["rng_seed"]))
    goal_keep = None
    if float(a["goal_drop"]) > 0:
        goal_keep = 1.0 - torch.bernoulli(torch.ones(a["goal"].shape) * float(a["goal_drop"]))
    masks = draw_dropout_masks(m.inner_model, a["action"].shape[0], a["action"].shape[1], "cpu")
    return cfg, sd, a, goal_keep, masks


@pytest.mark.parametrize("name", ["loss_dropout_K256", "loss_dropout_B256"])
def test_dropout_masks_replay_the_reference_generator(name):
    """Golden = the UNMODIFIED reference in training mode with dropout under torch.manual_seed; the oracle fed with the
    product's re-drawn masks must give the same loss and gradients (same masks <=> same generator consumption)."""
    from conftest import to_oracle_cfg
    cfg, sd, a, goal_keep, masks = _replay_training_draws(name)
    assert masks is not None and (masks["attn"][0] is not None)
    kw = dict(drop_masks=masks)
    if goal_keep is not None:
        kw["goal_keep"] = goal_keep
    loss, grads = O.loss_and_grads(sd, to_oracle_cfg(cfg), a["state"], a["action"], a["goal"], a["noise"].clone(), a["sigma"], **kw)
    torch.testing.assert_close(loss, a["loss"], rtol=1e-5, atol=1e-7)
    for n in [str(x) for x in a["grad_names"]]:
        flat = grads[n].reshape(-1)
        got = flat if flat.numel() <= 4096 else flat[::97][:4096]
        want = a["grad::" + n]
        torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-6 * float(want.abs().max()) + 2e-8)


@pytest.mark.parametrize("name", ["ckpt_push", "ckpt_kitchen2"])
def test_oracle_reproduces_trained_checkpoint_fixture(name):
    """The checkpoint fixtures (trained weights + the reference's outputs) are self-consistent: the oracle run on the
    stored weights gives the stored outputs, with or without /root/reference."""
    from conftest import load_checkpoint_golden
    cfg, meta, a, sd = load_checkpoint_golden(name)
    oc = to_oracle_cfg(cfg)
    with torch.no_grad():
        out = O.denoiser_forward(sd, oc, a["state"], a["action"], a["goal"], a["sigma"])
        ddim = O.sample_ddim(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_3"])
        anc = O.sample_euler_ancestral(sd, oc, a["state"], a["x_t"], a["goal"], a["sigmas_3"], noise=a["noise_3"])
        f16 = O.faithful16_denoiser_forward(sd, oc, a["state"], a["action"], a["goal"], a["sigma"])
    for got, key in ((out, "out"), (ddim, "ddim_3"), (anc, "euler_ancestral_3")):   # fp32 round-off: ATen blocks its GEMMs
        torch.testing.assert_close(got, a[key], rtol=1e-5, atol=2e-6)               # by thread count
    # the 16-bit-faithful oracle is a perturbation of the fp32 one of the size 16-bit operands cause, no more
    err = (f16 - a["out"]).abs().max()
    assert 1e-6 < float(err) < 3e-2
