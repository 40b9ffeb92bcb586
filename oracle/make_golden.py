"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz with the REAL reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

Every fixture is produced by calling the unmodified reference modules
(GCDenoiser / DiffusionGPT / gc_sampling / ClassifierFreeSampleModel) on CPU in
fp32 with weights from ``beso_b200.synth.synthetic_state_dict`` (regenerated
from the seed at test time; a float64 checksum of the weights is stored so a
drift of the generator is caught).  The reference cannot travel to the GPU box;
these files are what pins parity there.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from beso_b200.config import ModelConfig, K256, B256, T16          # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402
from oracle import ref_import                                      # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SMALL_KITCHEN = ModelConfig(obs_dim=30, act_dim=9, window=4, goal_len=2, d=360, n_layers=2, n_heads=6)
SMALL_PUSH = ModelConfig(obs_dim=10, act_dim=2, window=5, goal_len=1, d=240, n_layers=2, n_heads=12)
MLP_HEAD = ModelConfig(obs_dim=12, act_dim=3, window=3, goal_len=1, d=64, n_layers=1, n_heads=2, linear_output=False)
NO_GOAL = ModelConfig(obs_dim=12, act_dim=3, window=3, goal_len=2, d=64, n_layers=1, n_heads=2, goal_conditioned=False)

FWD_CASES = [  # name, cfg, weight seed, batch, t (None = W)
    ("fwd_K256", K256, 1, 8, None),
    ("fwd_K256_t1", K256, 1, 3, 1),
    ("fwd_K256_t4", K256, 1, 3, 4),
    ("fwd_T16", T16, 2, 8, None),
    ("fwd_B256", B256, 3, 8, None),
    ("fwd_small_kitchen", SMALL_KITCHEN, 4, 5, None),
    ("fwd_small_push", SMALL_PUSH, 5, 5, None),
    ("fwd_mlp_head", MLP_HEAD, 6, 4, None),
]


def cfg_dict(cfg: ModelConfig):
    return {k: getattr(cfg, k) for k in ("obs_dim", "act_dim", "window", "goal_len", "d", "n_layers",
                                        "n_heads", "sigma_data", "linear_output", "goal_conditioned")}


def weight_checksum(sd) -> float:
    return float(sum(v.double().sum().item() for k, v in sd.items() if not k.endswith("attn.mask")))


def build(ns, cfg, seed):
    m = ref_import.make_reference_model(ns, cfg)
    sd = synthetic_state_dict(cfg, seed)
    missing = m.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return m, sd


def save(name, cfg, seed, sd, **arrays):
    os.makedirs(OUT, exist_ok=True)
    meta = dict(cfg_dict(cfg), weight_seed=seed, weight_checksum=weight_checksum(sd),
                torch=torch.__version__)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array(repr(meta)),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                           for k, v in arrays.items()})
    print("wrote", name, {k: tuple(np.shape(v)) for k, v in arrays.items()})


@torch.no_grad()
def gen_forward(ns):
    for name, cfg, seed, B, t in FWD_CASES:
        m, sd = build(ns, cfg, seed)
        x = synthetic_inputs(cfg, B, seed=100 + seed, t=t)
        out = m(x["state"], x["action"], x["goal"], x["sigma"])
        out_u = m(x["state"], x["action"], x["goal"], x["sigma"], uncond=True)
        inner = m.inner_model(x["state"], x["action"], x["goal"], x["sigma"])
        save(name, cfg, seed, sd, state=x["state"], action=x["action"], goal=x["goal"], sigma=x["sigma"],
             out=out, out_uncond=out_u, inner=inner)
    # not goal conditioned (score_gpts.py:143-144,317-320,336-337); the reference indexes
    # pos_emb[:, goal_seq_len:] with the ORIGINAL goal_seq_len attribute overwritten to 0.
    cfg, seed = NO_GOAL, 7
    m, sd = build(ns, cfg, seed)
    x = synthetic_inputs(cfg, 4, seed=107)
    out = m(x["state"], x["action"], x["goal"], x["sigma"])
    save("fwd_no_goal", cfg, seed, sd, state=x["state"], action=x["action"], goal=x["goal"],
         sigma=x["sigma"], out=out)


@torch.no_grad()
def gen_samplers(ns):
    gs = ns.gc_sampling
    cfg, seed, B = K256, 1, 4
    m, sd = build(ns, cfg, seed)
    x = synthetic_inputs(cfg, B, seed=201)
    x_t = x["noise"] * 1.0
    arrays = dict(state=x["state"], goal=x["goal"], x_t=x_t)
    for n in (1, 3, 5):
        sig = gs.get_sigmas_exponential(n, 0.005, 1.0)
        arrays[f"sigmas_{n}"] = sig
        trace = []
        arrays[f"ddim_{n}"] = gs.sample_ddim(m, x["state"], x_t, x["goal"], sig, disable=True,
                                             callback=lambda d: trace.append(d["denoised"].clone()))
        if n == 5:
            arrays["ddim_5_denoised_trace"] = torch.stack(trace)
        arrays[f"euler_{n}"] = gs.sample_euler(m, x["state"], x_t, x["goal"], sig, disable=True)
        arrays[f"heun_{n}"] = gs.sample_heun(m, x["state"], x_t, x["goal"], sig, disable=True)
    sigk = gs.get_sigmas_karras(4, 0.005, 1.0, 5.0)
    arrays["sigmas_karras_4"] = sigk
    arrays["heun_karras_4"] = gs.sample_heun(m, x["state"], x_t, x["goal"], sigk, disable=True)
    # classifier-free guidance wrapper (classifier_free_sampler.py:12-52)
    CF = ns.classifier_free_sampler.ClassifierFreeSampleModel
    sig = gs.get_sigmas_exponential(4, 0.005, 1.0)
    arrays["sigmas_cfg_4"] = sig
    for lam in (0.0, 1.0, 1.5, 2.0):
        w = CF(m, cond_lambda=lam)
        tag = str(lam).replace(".", "p")
        arrays[f"cfg_fwd_{tag}"] = w(x["state"], x["action"], x["goal"], x["sigma"])
        arrays[f"cfg_heun4_{tag}"] = gs.sample_heun(w, x["state"], x_t, x["goal"], sig, disable=True)
        arrays[f"cfg_ddim4_{tag}"] = gs.sample_ddim(w, x["state"], x_t, x["goal"], sig, disable=True)
    arrays["action"] = x["action"]
    arrays["sigma"] = x["sigma"]
    save("samplers_K256", cfg, seed, sd, **arrays)


def gen_ancestral(ns):
    """sample_euler_ancestral of the real reference under a fixed generator state, plus the randn_like draws
    it consumed (re-drawn from the same state: the model itself uses no RNG in eval mode)."""
    gs = ns.gc_sampling
    cfg, seed, B = K256, 1, 4
    m, sd = build(ns, cfg, seed)
    x = synthetic_inputs(cfg, B, seed=201)
    x_t = x["noise"] * 1.0
    arrays = dict(state=x["state"], goal=x["goal"], x_t=x_t)
    for n in (1, 3, 5):
        sig = gs.get_sigmas_exponential(n, 0.005, 1.0)
        arrays[f"sigmas_{n}"] = sig
        torch.manual_seed(7000 + n)
        arrays[f"euler_ancestral_{n}"] = gs.sample_euler_ancestral(m, x["state"], x_t, x["goal"], sig, disable=True)
        torch.manual_seed(7000 + n)
        noise = torch.zeros((n,) + tuple(x_t.shape))
        for i in range(n):
            down, _ = gs.get_ancestral_step(sig[i], sig[i + 1])
            if down > 0:
                noise[i] = torch.randn_like(x_t)
        arrays[f"noise_{n}"] = noise
    sigk = gs.get_sigmas_karras(4, 0.005, 1.0, 5.0)
    arrays["sigmas_karras_4"] = sigk
    torch.manual_seed(7104)
    arrays["euler_ancestral_karras_4"] = gs.sample_euler_ancestral(m, x["state"], x_t, x["goal"], sigk, disable=True)
    torch.manual_seed(7104)
    noise = torch.zeros((4,) + tuple(x_t.shape))
    for i in range(4):
        down, _ = gs.get_ancestral_step(sigk[i], sigk[i + 1])
        if down > 0:
            noise[i] = torch.randn_like(x_t)
    arrays["noise_karras_4"] = noise
    # DPM-Solver++(2M) (deterministic), same inputs
    for n in (1, 3, 5):
        arrays[f"dpmpp_2m_{n}"] = gs.sample_dpmpp_2m(m, x["state"], x_t, x["goal"], arrays[f"sigmas_{n}"], disable=True)
    arrays["dpmpp_2m_karras_4"] = gs.sample_dpmpp_2m(m, x["state"], x_t, x["goal"], sigk, disable=True)
    # linear multistep (order 4 is reached on the 4th step)
    sig6 = gs.get_sigmas_exponential(6, 0.005, 1.0)
    arrays["sigmas_6"] = sig6
    for tag, sig in (("3", arrays["sigmas_3"]), ("6", sig6), ("karras_4", sigk)):
        arrays[f"lms_{tag}"] = gs.sample_lms(m, x["state"], x_t, x["goal"], sig, disable=True)
    # second-order single-step samplers (dpm_2, dpm_2_ancestral, dpmpp_2s, dpmpp_2s_ancestral)
    for tag, sig in (("3", arrays["sigmas_3"]), ("5", arrays["sigmas_5"]), ("karras_4", sigk)):
        n = len(sig) - 1
        torch.manual_seed(7200)                    # sample_dpm_2 draws (and discards) eps every step
        arrays[f"dpm_2_{tag}"] = gs.sample_dpm_2(m, x["state"], x_t, x["goal"], sig, disable=True)
        arrays[f"dpmpp_2s_{tag}"] = gs.sample_dpmpp_2s(m, x["state"], x_t, x["goal"], sig, disable=True)
        torch.manual_seed(7300 + n)
        arrays[f"dpm_2_ancestral_{tag}"] = gs.sample_dpm_2_ancestral(m, x["state"], x_t, x["goal"], sig, disable=True)
        torch.manual_seed(7300 + n)
        nz = torch.zeros((n,) + tuple(x_t.shape))
        for i in range(n):
            down, _ = gs.get_ancestral_step(sig[i], sig[i + 1])
            if down != 0:
                nz[i] = torch.randn_like(x_t)
        arrays[f"noise_dpm2a_{tag}"] = nz
        torch.manual_seed(7400 + n)
        arrays[f"dpmpp_2s_ancestral_{tag}"] = gs.sample_dpmpp_2s_ancestral(m, x["state"], x_t, x["goal"], sig, disable=True)
        torch.manual_seed(7400 + n)
        arrays[f"noise_2sa_{tag}"] = torch.stack([torch.randn_like(x_t) for _ in range(n)])
    save("samplers_ancestral_K256", cfg, seed, sd, **arrays)


def gen_dpm_solver(ns):
    """sample_dpm_fast / sample_dpm_adaptive of the real reference (host-side step control), K256, B = 4."""
    gs = ns.gc_sampling
    cfg, seed, B = K256, 1, 4
    m, sd = build(ns, cfg, seed)
    x = synthetic_inputs(cfg, B, seed=301)
    x_t = x["noise"] * 1.0
    arrays = dict(state=x["state"], goal=x["goal"], x_t=x_t)
    for n in (3, 6, 8, 10):                                   # nfe % 3 == 0 and != 0, as BesoAgent passes len(sigmas)
        torch.manual_seed(7500 + n)
        arrays[f"dpm_fast_{n}"] = gs.sample_dpm_fast(m, x["state"], x_t, x["goal"], 0.005, 1.0, n, disable=True)
    torch.manual_seed(7600)
    arrays["dpm_fast_eta_9"] = gs.sample_dpm_fast(m, x["state"], x_t, x["goal"], 0.005, 1.0, 9, disable=True, eta=0.5)
    torch.manual_seed(7600)
    arrays["noise_eta_9"] = torch.stack([torch.randn_like(x_t) for _ in range(4)])
    for order in (2, 3):
        torch.manual_seed(7700 + order)
        out, info = gs.sample_dpm_adaptive(m, x["state"], x_t, x["goal"], 0.005, 1.0, disable=True, order=order,
                                           return_info=True)
        arrays[f"dpm_adaptive_{order}"] = out
        arrays[f"dpm_adaptive_{order}_info"] = np.array([info[k] for k in ("steps", "nfe", "n_accept", "n_reject")])
    save("samplers_dpm_solver_K256", cfg, seed, sd, **arrays)


def gen_schedules(ns):
    gs = ns.gc_sampling
    arrays = {}
    for n in (1, 3, 10, 50):
        arrays[f"exponential_{n}"] = gs.get_sigmas_exponential(n, 0.005, 1.0)
        arrays[f"karras_{n}"] = gs.get_sigmas_karras(n, 0.005, 1.0, 5.0)
        arrays[f"linear_{n}"] = gs.get_sigmas_linear(n, 0.005, 1.0)
        arrays[f"vp_{n}"] = gs.get_sigmas_vp(n)
        if n > 1:
            arrays[f"ve_{n}"] = gs.get_sigmas_ve(n, 0.005, 1.0)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **{k: v.numpy() for k, v in arrays.items()})
    print("wrote schedules", list(arrays))


def gen_loss(ns):
    for name, cfg, seed, B in (("loss_B256", B256, 3, 16), ("loss_K256", K256, 1, 8)):
        m, sd = build(ns, cfg, seed)
        m.train()                      # beso_agent.py:229-230; dropout p = 0, goal_drop = 0
        m.training = True
        x = synthetic_inputs(cfg, B, seed=300 + seed, sigma_min=0.05)
        loss = m.loss(x["state"], x["clean"], x["goal"], x["noise"].clone(), x["sigma"])
        m.zero_grad()
        loss.backward()
        arrays = dict(state=x["state"], action=x["clean"], goal=x["goal"], noise=x["noise"],
                      sigma=x["sigma"], loss=loss.detach())
        norms, names = [], []
        for n, p in m.named_parameters():
            g = p.grad
            names.append(n)
            norms.append(g.double().norm().item())
            flat = g.reshape(-1)
            # full gradient for small tensors, a strided sample of the big ones
            arrays["grad::" + n] = flat if flat.numel() <= 4096 else flat[::97][:4096]
        arrays["grad_norms"] = np.array(norms)
        arrays["grad_names"] = np.array(names)
        loss_last = m.loss(x["state"], x["clean"], x["goal"], x["noise"].clone(), x["sigma"],
                           pred_last_action_only=True)
        arrays["loss_pred_last"] = loss_last.detach()
        save(name, cfg, seed, sd, **arrays)


DROPOUT_CASES = (("loss_dropout_K256", K256, 1, 8, dict(attn_pdrop=0.3, resid_pdrop=0.0, goal_drop=0.0)),      # kitchen config
                 ("loss_dropout_B256", B256, 3, 16, dict(attn_pdrop=0.05, resid_pdrop=0.05, goal_drop=0.1)))   # block-push + CFG mask
DROPOUT_SEED = 1234


def gen_loss_dropout(ns):
    """GCDenoiser.loss in training mode WITH dropout (configs/franka_kitchen_main_config.yaml:56-57 attn_pdrop 0.3;
    configs/block_push_main_config.yaml:57-58 attn/resid 0.05) under torch.manual_seed(DROPOUT_SEED) on CPU: the
    reference draws its masks from the global generator in op order.  The tests re-draw them with
    beso_b200.training.draw_dropout_masks under the same seed."""
    for name, cfg, seed, B, kw in DROPOUT_CASES:
        m = ref_import.make_reference_model(ns, cfg, **kw)
        sd = synthetic_state_dict(cfg, seed)
        m.load_state_dict(sd, strict=True)
        m.train()
        m.training = True
        x = synthetic_inputs(cfg, B, seed=300 + seed, sigma_min=0.05)
        torch.manual_seed(DROPOUT_SEED)
        loss = m.loss(x["state"], x["clean"], x["goal"], x["noise"].clone(), x["sigma"])
        m.zero_grad()
        loss.backward()
        arrays = dict(state=x["state"], action=x["clean"], goal=x["goal"], noise=x["noise"], sigma=x["sigma"],
                      loss=loss.detach(), attn_pdrop=kw["attn_pdrop"], resid_pdrop=kw["resid_pdrop"],
                      goal_drop=kw["goal_drop"], rng_seed=DROPOUT_SEED)
        norms, names = [], []
        for n, p in m.named_parameters():
            g = p.grad
            names.append(n)
            norms.append(g.double().norm().item())
            flat = g.reshape(-1)
            arrays["grad::" + n] = flat if flat.numel() <= 4096 else flat[::97][:4096]
        arrays["grad_norms"] = np.array(norms)
        arrays["grad_names"] = np.array(names)
        save(name, cfg, seed, sd, **arrays)


CKPT_CASES = (  # name, shipped checkpoint dir, config, layers kept
    ("ckpt_push", "trained_models/block_push/c_beso_1", "BLOCKPUSH_CKPT", None),
    ("ckpt_kitchen2", "trained_models/kitchen/c_beso_1", "KITCHEN_CKPT", 2),
)


def gen_real_checkpoints(ns):
    """TRAINED weights (the reference's shipped EMA checkpoints, trained_models/*/c_beso_1/model_state_dict.pth) through
    the unmodified reference: forward, unconditional forward, 3-step DDIM and 3-step Euler-ancestral.  So that the
    weights can travel as a fixture they are rounded to fp16 (stored as fp16, widened back to fp32 before the reference
    and every implementation under test load them: the same values everywhere); the kitchen model (9.4 M parameters)
    additionally keeps only its first two of six blocks.  Trained weights have the heavy-tailed statistics that make
    16-bit operand error 5x larger than on N(0, 0.02) initialisations."""
    import dataclasses
    from beso_b200 import config as C
    gs = ns.gc_sampling
    for name, rel, cfg_name, keep_layers in CKPT_CASES:
        cfg = getattr(C, cfg_name)
        sd = torch.load(os.path.join(ref_import.REF_ROOT, rel, "model_state_dict.pth"), map_location="cpu")
        if keep_layers is not None:
            cfg = dataclasses.replace(cfg, n_layers=keep_layers)
            sd = {k: v for k, v in sd.items() if not (k.startswith("inner_model.blocks.") and int(k.split(".")[2]) >= keep_layers)}
        sd = {k: (v.to(torch.float16).to(torch.float32) if not k.endswith("attn.mask") else v.float()) for k, v in sd.items()}
        m = ref_import.make_reference_model(ns, cfg)
        m.load_state_dict(sd, strict=True)
        m.eval()
        x = synthetic_inputs(cfg, 6, seed=900)
        with torch.no_grad():
            out = m(x["state"], x["action"], x["goal"], x["sigma"])
            out_u = m(x["state"], x["action"], x["goal"], x["sigma"], uncond=True)
            sig = gs.get_sigmas_exponential(3, 0.05, 1.0)
            ddim = gs.sample_ddim(m, x["state"], x["noise"], x["goal"], sig, disable=True)
            torch.manual_seed(901)
            noise = torch.stack([torch.randn_like(x["noise"]) for _ in range(3)])
            torch.manual_seed(901)
            anc = gs.sample_euler_ancestral(m, x["state"], x["noise"], x["goal"], sig, disable=True)
        arrays = dict(state=x["state"], action=x["action"], goal=x["goal"], sigma=x["sigma"], x_t=x["noise"], out=out,
                      out_uncond=out_u, sigmas_3=sig, ddim_3=ddim, noise_3=noise, euler_ancestral_3=anc)
        meta = dict(cfg_dict(cfg), checkpoint=rel, kept_layers=keep_layers, torch=torch.__version__)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array(repr(meta)),
                            **{k: v.numpy() for k, v in arrays.items()},
                            **{"w::" + k: v.to(torch.float16).numpy() for k, v in sd.items() if not k.endswith("attn.mask")})
        print("wrote", name, {k: tuple(v.shape) for k, v in arrays.items()}, f"{sum(v.numel() for v in sd.values()) / 1e6:.2f} M weights")


WINDOW_MODES = {
    "plain": dict(window=5),
    "future": dict(window=5, future_conditional=True, min_future_sep=1, future_seq_len=2),
    "tail": dict(window=4, future_conditional=True, min_future_sep=0, future_seq_len=3, only_sample_tail=True),
    "seq_end": dict(window=4, future_conditional=True, min_future_sep=0, future_seq_len=1, only_sample_seq_end=True),
}


def window_fixture_data():
    """Padded toy trajectories with ragged lengths (one shorter than every window, one exactly a window long)."""
    rs = np.random.RandomState(11)
    lens = np.array([12, 3, 5, 9, 20, 4, 7], dtype=np.int64)
    obs = rs.randn(len(lens), 20, 6).astype(np.float32)
    act = rs.randn(len(lens), 20, 3).astype(np.float32)
    for i, T in enumerate(lens):           # zero padding like the reference's padded datasets
        obs[i, T:] = 0
        act[i, T:] = 0
    return obs, act, lens


def reference_window_dataset(tl, obs, act, lens, **kw):
    """The unmodified TrajectorySlicerDataset over a minimal TrajectoryDataset of the padded arrays."""
    class Padded(tl.TrajectoryDataset):
        def __init__(self):
            self.obs, self.act = torch.from_numpy(obs), torch.from_numpy(act)
            self.mask = torch.from_numpy((np.arange(obs.shape[1])[None] < lens[:, None]).astype(np.float32))

        def __len__(self):
            return len(lens)

        def __getitem__(self, i):
            return self.obs[i], self.act[i], self.mask[i]

        def get_seq_length(self, i):
            return int(lens[i])

        def get_all_actions(self):
            return torch.cat([self.act[i, :int(T)] for i, T in enumerate(lens)])
    return tl.TrajectorySlicerDataset(Padded(), **kw)


def gen_windows():
    """tests/golden/windows.npz: batches of the reference's windowed dataset, every item in a seeded shuffled order."""
    tl = ref_import.load_trajectory_loader()
    obs, act, lens = window_fixture_data()
    arrays = dict(obs=obs, act=act, lens=lens)
    for mode, kw in WINDOW_MODES.items():
        ds = reference_window_dataset(tl, obs, act, lens, **kw)
        order = np.random.RandomState(5).permutation(len(ds))
        np.random.seed(1234)
        loader = torch.utils.data.DataLoader(torch.utils.data.Subset(ds, order.tolist()), batch_size=len(ds), shuffle=False)
        (b,) = list(loader)
        arrays[f"{mode}::order"] = order
        arrays[f"{mode}::slices"] = np.array(ds.slices, dtype=np.int64)
        for k, v in b.items():
            arrays[f"{mode}::{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "windows.npz"), **arrays)
    print("wrote windows.npz", {k: v.shape for k, v in arrays.items()})


def scaler_fixture_data():
    rs = np.random.RandomState(21)
    x = (rs.randn(12, 30, 6) * np.array([1, 5, 0.1, 2, 1, 3]) + np.array([0, 2, -1, 4, 0, 1])).astype(np.float32)
    y = (rs.randn(12, 30, 3) * np.array([2, 0.5, 1]) + np.array([1, 0, -3])).astype(np.float32)
    return x, y


def gen_scalers():
    """tests/golden/scalers.npz: statistics and method outputs of the reference's Scaler / MinMaxScaler."""
    import importlib.util
    path = os.path.join(ref_import.REF_ROOT, "beso", "networks", "scaler", "scaler_class.py")
    spec = importlib.util.spec_from_file_location("_beso_ref_scaler", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    x, y = scaler_fixture_data()
    xs, ys = torch.from_numpy(x[:4]), torch.from_numpy(y[:4])
    arrays = {}
    for cls in ("Scaler", "MinMaxScaler"):
        for on in (True, False):
            sc = getattr(ref, cls)(x, y, on, "cpu")
            tag = f"{cls}::{int(on)}::"
            for name in ("x_mean", "x_std", "x_max", "x_min", "y_min", "y_max", "y_bounds_tensor", "x_bounds_tensor"):
                arrays[tag + name] = getattr(sc, name).numpy()
            for fn, arg in (("scale_input", xs), ("scale_output", ys), ("inverse_scale_input", xs),
                            ("inverse_scale_output", ys), ("clip_action", ys * 3)):
                arrays[tag + fn] = getattr(sc, fn)(arg.clone()).numpy()
    np.savez_compressed(os.path.join(OUT, "scalers.npz"), **arrays)
    print("wrote scalers.npz", len(arrays), "arrays")


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    if len(sys.argv) > 1 and sys.argv[1] == "windows":         # dataset fixture: needs only trajectory_loader.py
        gen_windows()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "scalers":         # needs only scaler_class.py
        gen_scalers()
        return
    ns = ref_import.load()
    if len(sys.argv) > 1 and sys.argv[1] == "ancestral":       # only the fixture added after the first set
        gen_ancestral(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dpm_solver":
        gen_dpm_solver(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "real_ckpt":       # added in round 2
        gen_real_checkpoints(ns)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "loss_dropout":    # added in round 2
        gen_loss_dropout(ns)
        return
    gen_ancestral(ns)
    gen_dpm_solver(ns)
    gen_schedules(ns)
    gen_forward(ns)
    gen_samplers(ns)
    gen_loss(ns)
    gen_loss_dropout(ns)
    gen_real_checkpoints(ns)


if __name__ == "__main__":
    main()
