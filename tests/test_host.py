"""CPU: host logic, the C-ABI surface and the no-fallback guarantee (no GPU compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, golden_weights, load_golden
from beso_b200 import B256, K256, KITCHEN_CKPT, T16, _lib, sampling
from beso_b200.agent import BesoAgent
from beso_b200.cfg import ClassifierFreeSampleModel
from beso_b200.denoiser import DiffusionGPT, GCDenoiser, build_denoiser
from beso_b200.synth import synthetic_state_dict


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.lib()
    header = open(os.path.join(ROOT, "include", "beso_b200.h")).read()
    declared = set(re.findall(r"\b(beso_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.beso_abi_version() == 1


@pytest.mark.parametrize("cfg", [K256, B256, T16, KITCHEN_CKPT])
def test_param_table_matches_module(cfg):
    lib = _lib.lib()
    desc = _lib.ModelDesc.from_config(cfg)
    m = build_denoiser(cfg, "cpu")
    params = list(m.get_params())
    assert lib.beso_param_count(C.byref(desc)) == len(params)
    assert lib.beso_param_total(C.byref(desc)) == sum(p.numel() for p in params) == cfg.n_params()
    for i, p in enumerate(params):
        assert lib.beso_param_numel(C.byref(desc), i) == p.numel()
    assert [n for n, _ in m.named_parameters()] == [n for n, _ in cfg.param_shapes()]


def test_invalid_descriptions_are_rejected():
    lib = _lib.lib()
    bad = _lib.ModelDesc(60, 9, 10, 2, 250, 4, 4, 1, 1, 0.5)      # d % n_heads != 0
    assert lib.beso_param_count(C.byref(bad)) == -1
    assert b"invalid" in lib.beso_last_error()
    h = C.c_void_p()
    assert lib.beso_plan_create(C.byref(bad), 0, C.byref(h)) == -1


def test_state_dict_schema_matches_reference_fixture_weights():
    cfg, meta, _ = load_golden("fwd_K256")
    sd = golden_weights(cfg, meta)
    m = build_denoiser(cfg, "cpu")
    assert list(m.state_dict().keys()) == list(sd.keys())
    m.load_state_dict(sd, strict=True)
    assert m.state_dict()["inner_model.blocks.0.attn.mask"].shape == (1, 1, 23, 23)
    assert m.inner_model.pos_emb.shape == (1, 13, 256)            # seq_size = G + W + 1, last row unused


def test_state_dict_schema_matches_live_reference():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("/root/reference not present")
    ns = ref_import.load()
    for cfg in (K256, KITCHEN_CKPT):
        ref = ref_import.make_reference_model(ns, cfg)
        ours = build_denoiser(cfg, "cpu")
        assert [(k, tuple(v.shape)) for k, v in ref.state_dict().items()] == \
               [(k, tuple(v.shape)) for k, v in ours.state_dict().items()]
        assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    sd = torch.load(os.path.join(ref_import.REF_ROOT, "trained_models/kitchen/c_beso_1/model_state_dict.pth"),
                    map_location="cpu")
    build_denoiser(KITCHEN_CKPT, "cpu").load_state_dict(sd, strict=True)


def test_hydra_style_instantiation():
    spec = dict(_target_="beso_b200.denoiser.DiffusionGPT", state_dim=16, device="cpu", goal_conditioned=True,
                action_dim=2, embed_dim=64, embed_pdrob=0, attn_pdrop=0.1, resid_pdrop=0.1, n_layers=2, n_heads=4,
                goal_seq_len=1, obs_seq_len=5, sigma_vocab_size=3, time_embedding_fn={"x": 1}, goal_drop=0.1,
                linear_output=True)
    m = GCDenoiser(spec, sigma_data=0.5)
    assert isinstance(m.inner_model, DiffusionGPT) and m.inner_model.block_size == 12
    assert m.config.sigma_data == 0.5 and m.config.n_tokens() == 12
    m.training = True                         # the agent assigns the attribute directly (beso_agent.py:230)
    m.min_action = 0.0                        # arbitrary attribute set (beso_agent.py:116-117)
    assert len(list(m.get_params())) == len(list(m.parameters()))


def test_no_cpu_fallback():
    m = build_denoiser(K256, "cpu")
    with pytest.raises(_lib.BesoLibraryError):
        m(torch.zeros(1, 10, 60), torch.zeros(1, 10, 9), torch.zeros(1, 2, 60), torch.ones(1))
    with pytest.raises(_lib.BesoLibraryError):
        sampling.sample_ddim(ClassifierFreeSampleModel(m, 2.0), torch.zeros(1, 10, 60), torch.zeros(1, 10, 9),
                             torch.zeros(1, 2, 60), sampling.get_sigmas_exponential(3, 0.005, 1.0))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "beso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("beso_oracle", "oracle") or f == "__none__", f
                assert "/root/reference" not in src, f


def test_schedules_match_reference_fixture():
    z = np.load(os.path.join(ROOT, "tests", "golden", "schedules.npz"))
    for n in (1, 3, 10, 50):
        np.testing.assert_array_equal(sampling.get_sigmas_exponential(n, 0.005, 1.0).numpy(), z[f"exponential_{n}"])
        np.testing.assert_array_equal(sampling.get_sigmas_karras(n, 0.005, 1.0, 5.0).numpy(), z[f"karras_{n}"])
        np.testing.assert_array_equal(sampling.get_sigmas_linear(n, 0.005, 1.0).numpy(), z[f"linear_{n}"])
        np.testing.assert_array_equal(sampling.get_sigmas_vp(n).numpy(), z[f"vp_{n}"])
        if n > 1:
            np.testing.assert_array_equal(sampling.get_sigmas_ve(n, 0.005, 1.0).numpy(), z[f"ve_{n}"])


def test_ddim_coefficients_last_step_is_exact_replacement():
    c = sampling.ddim_coefficients(sampling.get_sigmas_exponential(5, 0.005, 1.0))
    assert c.shape == (5, 2)
    assert c[-1, 0] == 0.0 and c[-1, 1] == -1.0          # x <- 0 * x - (-1) * denoised


def test_eval_counts():
    s = sampling.get_sigmas_exponential(10, 0.005, 1.0)
    assert sampling.n_model_evals("ddim", s) == 10 and sampling.n_model_evals("heun", s) == 19


def test_python_loop_samplers_against_oracle_with_a_torch_model():
    """The step-by-step fallback loops (used with callbacks / churn) follow the reference update
    rules: run them around the oracle's forward and compare with the reference fixtures."""
    from conftest import to_oracle_cfg
    from oracle import beso_oracle as O
    cfg, meta, a = load_golden("samplers_K256")
    sd, oc = O.as_module_params(golden_weights(cfg, meta)), to_oracle_cfg(cfg)
    model = lambda s, x, g, sig, **kw: O.denoiser_forward(sd, oc, s, x, g, sig, **kw)   # noqa: E731
    seen = []
    for name in ("ddim", "euler", "heun"):
        got = sampling.SAMPLERS[name](model, a["state"], a["x_t"], a["goal"], a["sigmas_5"],
                                      callback=lambda d: seen.append(d["i"]))
        torch.testing.assert_close(got, a[f"{name}_5"], rtol=1e-5, atol=1e-6)
    assert seen[:5] == [0, 1, 2, 3, 4]


def test_agent_dispatch_and_errors():
    m = build_denoiser(K256, "cpu")
    agent = BesoAgent(m, device="cpu", window_size=10)
    with pytest.raises(ValueError):
        agent.sample_loop(torch.ones(3), torch.zeros(1, 10, 9), torch.zeros(1, 10, 60), torch.zeros(1, 2, 60), "nope")
    with pytest.raises(ValueError):
        agent.get_noise_schedule(3, "nope")
    with pytest.raises(KeyError):               # non-empty extra_args must hold both keys
        agent.sample_loop(torch.ones(3), torch.zeros(1, 10, 9), torch.zeros(1, 10, 60), torch.zeros(1, 2, 60),
                          "ddim", {"s_churn": 1})
    assert torch.equal(agent.get_noise_schedule(3, "exponential"), sampling.get_sigmas_exponential(3, 0.005, 1.0))


def test_ancestral_coefficients_follow_the_reference_formula():
    """(sigma_down, sigma_up) per step = get_ancestral_step (gc_sampling.py:108-114) in fp32 tensor ops."""
    sig = sampling.get_sigmas_exponential(5, 0.005, 1.0)
    coef = sampling.ancestral_coefficients(sig)
    assert coef.shape == (5, 2) and coef.dtype == torch.float32
    for i in range(5):
        s_from, s_to = sig[i], sig[i + 1]
        up = min(s_to, (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5)
        down = (s_to ** 2 - up ** 2) ** 0.5
        assert float(coef[i, 0]) == float(down) and float(coef[i, 1]) == float(up)
    assert float(coef[-1, 0]) == 0.0 and float(coef[-1, 1]) == 0.0      # last step: onto sigma = 0, no noise
    assert "euler_ancestral" in sampling.SAMPLERS and _lib.SAMPLER_IDS["euler_ancestral"] == 3


def test_euler_ancestral_refuses_cpu_models_like_every_other_sampler():
    m = build_denoiser(K256, "cpu")
    x = torch.zeros(2, K256.window, K256.act_dim)
    s = torch.zeros(2, K256.window, K256.obs_dim)
    g = torch.zeros(2, K256.goal_len, K256.obs_dim)
    with pytest.raises(_lib.BesoLibraryError):
        sampling.sample_euler_ancestral(m, s, x, g, sampling.get_sigmas_exponential(3, 0.005, 1.0))


def test_dpmpp_2m_coefficients_follow_the_reference_formula():
    sig = sampling.get_sigmas_exponential(4, 0.005, 1.0)
    coef = sampling.dpmpp_2m_coefficients(sig)
    assert coef.shape == (4, 4)
    assert float(coef[0, 2]) == 0.0 and float(coef[0, 3]) == 0.0          # first step: first order
    assert float(coef[-1, 2]) == 0.0 and float(coef[-1, 3]) == 0.0        # step onto sigma = 0: first order
    t = sig[:-1].log().neg()
    h1, h0 = (sig[2].log().neg() - t[1]), (t[1] - t[0])
    r = h0 / h1
    assert float(coef[1, 2]) == float(1 + 1 / (2 * r)) and float(coef[1, 3]) == float(1 / (2 * r))
    assert torch.equal(coef[:, :2], sampling.ddim_coefficients(sig))
    assert _lib.SAMPLER_IDS["dpmpp_2m"] == 4 and "dpmpp_2m" in sampling.SAMPLERS


def test_sample_density_matches_reference_formulas():
    """make_sample_density (beso_agent.py:540-578) -> utils.rand_log_logistic / rand_log_normal (utils.py:170-190):
    same torch calls, so the same generator state gives the same noise levels."""
    import math
    m = build_denoiser(K256, "cpu")
    agent = BesoAgent(m, device="cpu", sigma_min=0.005, sigma_max=1.0)
    agent.sigma_sample_density_type = "loglogistic"
    agent.sigma_sample_density_mean, agent.sigma_sample_density_std = -1.2, 1.2
    torch.manual_seed(3)
    got = agent.make_sample_density()(shape=(1000,), device="cpu")
    torch.manual_seed(3)
    loc, scale = math.log(m.sigma_data), 0.5
    lo, hi = torch.as_tensor(0.005, dtype=torch.float64), torch.as_tensor(1.0, dtype=torch.float64)
    min_cdf, max_cdf = lo.log().sub(loc).div(scale).sigmoid(), hi.log().sub(loc).div(scale).sigmoid()
    u = torch.rand((1000,), dtype=torch.float64) * (max_cdf - min_cdf) + min_cdf
    want = u.logit().mul(scale).add(loc).exp().to(torch.float32)
    assert torch.equal(got, want) and float(got.min()) >= 0.005 and float(got.max()) <= 1.0
    agent.sigma_sample_density_type = "lognormal"
    torch.manual_seed(4)
    got = agent.make_sample_density()(shape=(64,), device="cpu")
    torch.manual_seed(4)
    assert torch.equal(got, (torch.randn((64,)) * 1.2 - 1.2).exp())
    agent.sigma_sample_density_type = "nope"
    with pytest.raises(ValueError):
        agent.make_sample_density()


def test_two_stage_coefficient_rows():
    """[sigma_b, a1, b1, a2, b2, c2, su, 0] per step: last step onto sigma = 0 is a single-stage Euler step, the
    ancestral variants carry sigma_up, and the agent dispatches the reference's sampler names."""
    sig = sampling.get_sigmas_exponential(4, 0.005, 1.0)
    for kind in ("dpm_2", "dpmpp_2s", "dpm_2_ancestral", "dpmpp_2s_ancestral"):
        coef = sampling.two_stage_coefficients(kind, sig)
        assert coef.shape == (4, 8)
        assert float(coef[-1, 0]) == 0.0                                   # last step: one evaluation
        assert all(float(coef[i, 0]) > 0.0 for i in range(3))              # two evaluations otherwise
        assert (coef[:, 6] != 0).any().item() == kind.endswith("ancestral")
        # the last step lands on the denoised sample: x = 0 * x + 1 * D
        assert float(coef[-1, 1]) == pytest.approx(0.0, abs=1e-6) and float(coef[-1, 2]) == pytest.approx(1.0, abs=1e-6)
    mid = float(sampling.two_stage_coefficients("dpm_2", sig)[0, 0])
    assert mid == pytest.approx((float(sig[0]) * float(sig[1])) ** 0.5, rel=1e-6)   # log-space midpoint
    with pytest.raises(ValueError):
        sampling.two_stage_coefficients("nope", sig)
    for name in ("dpm", "ancestral", "dpmpp_2s", "dpmpp_2s_ancestral", "dpmpp_2m", "euler_ancestral"):
        assert name in sampling.SAMPLERS
    s10 = sampling.get_sigmas_exponential(10, 0.005, 1.0)
    assert sampling.n_model_evals("dpm", s10) == 19 and sampling.n_model_evals("ancestral", s10) == 19


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use
    it.  The package, the C/CUDA sources and the tools must not mention it as an import."""
    import pathlib
    import re
    root = pathlib.Path(__file__).resolve().parents[1]
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|importlib\.import_module\(\s*['\"]oracle", re.M)
    offenders = []
    for sub in ("beso_b200", "tools"):
        for f in (root / sub).rglob("*.py"):
            if pat.search(f.read_text()):
                offenders.append(str(f.relative_to(root)))
    assert offenders == []
    for f in (root / "beso_b200" / "csrc").glob("*.cu*"):
        assert "oracle/" not in f.read_text()


def test_process_batch_follows_the_reference_rules():
    """base_agent.py:111-142: scaling, the zeroed block-push goal dims, return conventions, pre-scaled batches."""
    import numpy as np
    from beso_b200 import scaler as S
    rs = np.random.RandomState(0)
    x, y = rs.randn(40, 10).astype(np.float32) * 3 + 1, rs.randn(40, 2).astype(np.float32)
    sc = S.Scaler(x, y, True, "cpu")
    agent = BesoAgent(build_denoiser(K256, "cpu"), device="cpu", window_size=10, scaler=sc)
    batch = {"observation": torch.from_numpy(x[:6]).view(2, 3, 10), "goal_observation": torch.from_numpy(x[6:8]).view(2, 1, 10),
             "action": torch.from_numpy(y[:6]).view(2, 3, 2)}
    state, action, goal = agent.process_batch(batch, predict=False)
    assert torch.equal(state, sc.scale_input(batch["observation"])) and torch.equal(action, sc.scale_output(batch["action"]))
    want_goal = sc.scale_input(batch["goal_observation"])
    want_goal[..., [2, 5, 6, 7, 8, 9]] = 0
    assert torch.equal(goal, want_goal) and float(goal[..., [0, 1, 3, 4]].abs().min()) > 0
    no_action = {k: v for k, v in batch.items() if k != "action"}
    s2, g2, name = agent.process_batch(dict(no_action, goal_task_name="push"), predict=True)
    assert name == "push" and torch.equal(g2, want_goal) and torch.equal(s2, state)
    assert agent.process_batch(no_action, predict=True)[2] is None and len(agent.process_batch(no_action, predict=False)) == 2
    # pre-scaled batches (DeviceWindowDataset with the scaler fused) are only moved and masked
    pre = {"observation": state.clone(), "goal_observation": sc.scale_input(batch["goal_observation"]), "action": action.clone(),
           "scaled": True}
    s3, a3, g3 = agent.process_batch(pre, predict=False)
    assert torch.equal(s3, state) and torch.equal(a3, action) and torch.equal(g3, want_goal)
    # goals with another feature count are left alone
    agent60 = BesoAgent(build_denoiser(K256, "cpu"), device="cpu", window_size=10)
    g = torch.ones(1, 2, 60)
    assert torch.equal(agent60.process_batch({"observation": torch.ones(1, 10, 60), "goal_observation": g})[1], g)


def test_checkpoint_files_round_trip(tmp_path):
    """store_model_weights / load_pretrained_model (beso_agent.py:458-477): EMA weights in model_state_dict.pth, raw
    ones in non_ema_model_state_dict.pth, buffers kept, nothing swapped in the live model."""
    from beso_b200.synth import synthetic_state_dict
    cfg = K256
    m = build_denoiser(cfg, "cpu", state_dict=synthetic_state_dict(cfg, 3))
    ema = [p.detach() * 0.5 for p in m.get_params()]
    agent = BesoAgent(m, device="cpu", window_size=cfg.window, use_ema=True, ema_params=ema)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    agent.store_model_weights(str(tmp_path))
    raw = torch.load(tmp_path / "non_ema_model_state_dict.pth")
    avg = torch.load(tmp_path / "model_state_dict.pth")
    assert list(raw) == list(before) == list(avg)
    for k, v in before.items():
        assert torch.equal(raw[k], v) and torch.equal(m.state_dict()[k], v)
        assert torch.equal(avg[k], v if k.endswith("attn.mask") else v * 0.5)
    other = BesoAgent(build_denoiser(cfg, "cpu"), device="cpu", window_size=cfg.window)
    other.load_pretrained_model(str(tmp_path))
    for k, v in other.model.state_dict().items():
        assert torch.equal(v, avg[k])
    # scaler hand-over (beso_agent.py:106-117)
    import numpy as np
    from beso_b200 import scaler as S
    sc = S.Scaler(np.random.RandomState(0).randn(20, 60).astype(np.float32), np.random.RandomState(1).randn(20, 9).astype(np.float32),
                  True, "cpu")
    other.get_scaler(sc)
    other.set_bounds(sc)
    assert other.scaler is sc and other.model.min_action.shape == (9,) and bool((other.model.max_action > other.model.min_action).all())


@pytest.mark.skipif(not os.path.isdir("/root/reference/trained_models/kitchen/c_beso_1"), reason="shipped checkpoints not present")
def test_load_shipped_checkpoint():
    from beso_b200.config import KITCHEN_CKPT
    agent = BesoAgent(build_denoiser(KITCHEN_CKPT, "cpu"), device="cpu", window_size=KITCHEN_CKPT.window)
    agent.load_pretrained_model("/root/reference/trained_models/kitchen/c_beso_1")
    sd = torch.load("/root/reference/trained_models/kitchen/c_beso_1/model_state_dict.pth", map_location="cpu")
    assert all(torch.equal(v, sd[k]) for k, v in agent.model.state_dict().items())


def test_train_step_host_flow(monkeypatch):
    """The host side of BesoAgent.train_step without a GPU: batch processing, RNG draws, the data-parallel hook between
    gradient and optimiser, scheduler / EMA bookkeeping.  The fused loss + backward is stubbed."""
    import beso_b200.training as T
    from beso_b200.synth import synthetic_inputs
    agent = BesoAgent(build_denoiser(K256, "cpu"), device="cpu", window_size=10)
    seen = {}

    def fake(core, state, action, goal, noise, sigma, pred_last, goal_keep, dropout_masks=None, grad_sync=None):
        seen["args"] = (state.shape, action.shape, goal.shape, noise.clone(), sigma.clone(), pred_last, goal_keep)
        return torch.tensor(1.5), torch.arange(10.0)
    monkeypatch.setattr(T, "loss_and_flat_grad", fake)
    order = []
    agent.optimizer = type("O", (), {"step": lambda self, flat_grad=None: order.append(("opt", flat_grad.clone()))})()
    agent.lr_scheduler = type("S", (), {"step": lambda self: order.append(("sched",))})()
    agent.ema_helper = type("E", (), {"update": lambda self, p: order.append(("ema",))})()
    agent.update_ema_every_n_steps, agent.steps = 2, 0
    agent.sigma_sample_density_type, agent.sigma_sample_density_mean, agent.sigma_sample_density_std = "loglogistic", -1.2, 1.2
    agent.grad_sync = lambda flat: order.append(("sync",)) or flat.mul_(0.5)
    x = synthetic_inputs(K256, 4, seed=1)
    batch = {"observation": x["state"], "goal_observation": x["goal"], "action": x["clean"]}
    torch.manual_seed(5)
    assert agent.train_step(batch) == 1.5
    torch.manual_seed(5)                                     # the reference's draw order: noise first, then sigma
    noise = torch.randn_like(x["clean"])
    sigma = agent.make_sample_density()(shape=(4,), device="cpu")
    assert torch.equal(seen["args"][3], noise) and torch.equal(seen["args"][4], sigma)
    assert seen["args"][:3] == (x["state"].shape, x["clean"].shape, x["goal"].shape) and seen["args"][5] is False
    assert [o[0] for o in order] == ["sync", "opt", "sched"]           # EMA every 2nd step only
    assert torch.equal(order[1][1], torch.arange(10.0) * 0.5)           # the optimiser sees the all-reduced gradient
    agent.train_step(batch)
    assert [o[0] for o in order][3:] == ["sync", "opt", "sched", "ema"] and agent.steps == 2


def test_predict_host_flow(monkeypatch):
    """predict() without a GPU (sample loop stubbed): observation / action context deques, the re-denoised action
    context, scaler clip + inverse scaling (beso_agent.py:297-388)."""
    import numpy as np
    from beso_b200 import scaler as S
    from beso_b200 import T16
    cfg = T16
    rs = np.random.RandomState(0)
    sc = S.Scaler(rs.randn(50, cfg.obs_dim).astype(np.float32), rs.randn(50, cfg.act_dim).astype(np.float32), True, "cpu")
    agent = BesoAgent(build_denoiser(cfg, "cpu"), device="cpu", window_size=cfg.window, num_sampling_steps=3, scaler=sc)
    shapes = []

    def fake_loop(sigmas, x_t, state, goal, sampler_type, extra_args={}):
        shapes.append((tuple(x_t.shape), tuple(state.shape), tuple(goal.shape), sampler_type, len(sigmas)))
        return torch.full_like(x_t, 100.0)                    # far outside the action bounds: must be clipped
    monkeypatch.setattr(agent, "sample_loop", fake_loop)
    for step in range(cfg.window + 2):
        out = agent.predict({"observation": torch.randn(1, cfg.obs_dim), "goal_observation": torch.randn(cfg.goal_len, cfg.obs_dim)},
                            new_sampler_type="euler", extra_args={})
        t = min(step + 1, cfg.window)
        assert shapes[-1] == ((1, t, cfg.act_dim), (1, t, cfg.obs_dim), (1, cfg.goal_len, cfg.obs_dim), "euler", 4)
        clipped = sc.y_bounds_tensor[1, :].float() * 1.1
        assert out.shape == ((1, 1, cfg.act_dim) if step == 0 else (1, cfg.act_dim))   # the reference slices only when t > 1
        torch.testing.assert_close(out.reshape(1, -1), sc.inverse_scale_output(clipped.view(1, -1)))
        assert len(agent.action_context) == min(step + 1, cfg.window - 1)
    agent.reset()
    assert len(agent.obs_context) == 0 and len(agent.action_context) == 0


def test_library_carries_tcgen05_and_tma_code():
    """The fast path is tensor-core code, not a recompiled mma.sync kernel: the sm_100a SASS of the library holds
    tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM) and bulk async copies (UBLKCP), and only sm_100a code."""
    import shutil
    import subprocess
    from beso_b200 import _lib
    if shutil.which("cuobjdump") is None or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("UTCHMMA") > 100 and sass.count("LDTM") > 10 and sass.count("UBLKCP") > 10
    archs = set(__import__("re").findall(r"arch = (sm_\w+)", sass))
    assert archs == {"sm_100a"}, archs


def test_public_header_is_plain_c():
    """include/beso_b200.h is the FFI contract: it must parse as C99 (and C++) on its own, no CUDA or torch headers."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "beso_b200.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    includes = [l for l in open(hdr).read().splitlines() if l.lstrip().startswith("#include")]
    assert includes and all("cuda" not in l and "torch" not in l for l in includes), includes


def test_parameter_slot_cache_tracks_what_the_module_tree_walk_would():
    """GCDenoiser._params() avoids walking the module tree on every call (the walk cost more host time than a batch-1
    forward takes on the GPU).  It must stay equal to parameters() through every way the reference touches a model:
    in-place updates (optimizer, EMA copy), load_state_dict, dtype / device conversion, a re-assigned Parameter object."""
    import torch.nn as nn
    m = build_denoiser(K256, "cpu")

    def same():
        return [id(p) for p in m._params()] == [id(p) for p in m.inner_model.parameters()]
    assert same() and len(m._params()) == len(K256.param_shapes())
    fp0 = m._fingerprint()
    with torch.no_grad():
        m.inner_model.ln_f.weight.add_(1.0)                      # in-place update bumps the version counter
    assert same() and m._fingerprint() != fp0
    m.load_state_dict(m.state_dict())
    assert same()
    m.double()
    m.float()
    assert same()
    fp1 = m._fingerprint()
    m.inner_model.ln_f.weight = nn.Parameter(torch.ones_like(m.inner_model.ln_f.weight))   # new Parameter object
    assert same() and m._fingerprint() != fp1
