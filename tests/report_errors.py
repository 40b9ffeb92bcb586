"""Measured error of the tensor-core kernel's two arithmetic modes on the B200: against the reference goldens (random
init and TRAINED checkpoint weights) and, for the fp16 mode, against the 16-bit-faithful oracle (the reference with the
kernel's operand roundings, oracle/beso_oracle.py).  A checker script, not a pytest module (it lives under tests/
because only tests/ may import the oracle).  python tests/report_errors.py > profiles/r2_error_report.txt"""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_weights, load_checkpoint_golden, load_golden, to_oracle_cfg, with_masks  # noqa: E402
from beso_b200 import sampling                                   # noqa: E402
from beso_b200.denoiser import build_denoiser                    # noqa: E402
from oracle import beso_oracle as O                              # noqa: E402

dev = torch.device("cuda:0")


def line(tag, got, want):
    err = (got - want).abs()
    need = (err / (1e-3 * want.abs() + 1e-5)).max()
    print(f"{tag}: max|err| {float(err.max()):.2e} mean {float(err.mean()):.2e} max|ref| {float(want.abs().max()):.2f} "
          f"inside 1e-3/1e-5: {float((err <= 1e-5 + 1e-3 * want.abs()).float().mean()):.3f} worst err/tol {float(need):.1f}")
    return float(err.max())


def use(mode):
    """The layout hook is global and read at launch time: set it before every use of a model."""
    from beso_b200 import _lib
    _lib.lib().beso_debug_set_precise_layout({"precise128": 2}.get(mode, 0))


def models_for(cfg, sd, strict_masks=False):
    """precise = the precise mode as launched (small batches: stacked 64-row tiles), precise128 = its 128-row tile layout
    forced (embed_dim <= 256), fast = fp16 mode."""
    from conftest import build_for_mode
    out = {}
    for mode in ("precise", "precise128", "fast"):
        if mode == "precise128" and cfg.d > 256:
            continue
        m = build_for_mode(cfg, dev, mode=mode)
        m.load_state_dict(with_masks(m, sd) if strict_masks else sd, strict=True)
        m.eval()
        if mode == "fast" and not m.fast_supported():
            continue
        out[mode] = m
    return out


print("== forward goldens (random-init weights, N(0, 0.02)) ==")
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fwd_*.npz"))):
    name = os.path.basename(path)[:-4]
    cfg, meta, a = load_golden(name)
    sd = golden_weights(cfg, meta)
    g = {k: v.to(dev) for k, v in a.items() if isinstance(v, torch.Tensor)}
    for mode, m in models_for(cfg, sd).items():
        use(mode)
        kind = ("tensor-core fp16" if mode == "fast" else
                ("CUDA-core fp32" if not m.fast_supported() else ("tensor-core split, 128-row tiles" if mode == "precise128" else "tensor-core split, stacked tiles")))
        out = m(g["state"], g["action"], g["goal"], g["sigma"]).cpu()
        line(f"{name} [{mode}: {kind}] vs reference", out, a["out"])
        if mode == "fast":
            with torch.no_grad():
                f16 = O.faithful16_denoiser_forward(sd, to_oracle_cfg(cfg), a["state"], a["action"], a["goal"], a["sigma"])
            line(f"{name} [fast] vs 16-bit-faithful oracle", out, f16)
            line(f"{name} (16-bit-faithful oracle vs reference)", f16, a["out"])

print("== TRAINED checkpoint weights (trained_models/*/c_beso_1, fp16-rounded fixture) ==")
for name in ("ckpt_push", "ckpt_kitchen2"):
    cfg, meta, a, sd = load_checkpoint_golden(name)
    g = {k: v.to(dev) for k, v in a.items() if isinstance(v, torch.Tensor)}
    for mode, m in models_for(cfg, sd, strict_masks=True).items():
        use(mode)
        kind = ("tensor-core fp16" if mode == "fast" else
                ("CUDA-core fp32" if not m.fast_supported() else ("tensor-core split, 128-row tiles" if mode == "precise128" else "tensor-core split, stacked tiles")))
        out = m(g["state"], g["action"], g["goal"], g["sigma"]).cpu()
        line(f"{name} forward [{mode}: {kind}] vs reference", out, a["out"])
        got = sampling.sample_ddim(m, g["state"], g["x_t"], g["goal"], a["sigmas_3"]).cpu()
        line(f"{name} ddim_3 [{mode}] vs reference", got, a["ddim_3"])
        got = sampling.sample_euler_ancestral(m, g["state"], g["x_t"], g["goal"], a["sigmas_3"], noise=g["noise_3"]).cpu()
        line(f"{name} euler_ancestral_3 [{mode}] vs reference", got, a["euler_ancestral_3"])
        if mode == "fast":
            full = with_masks(m, sd)
            with torch.no_grad():
                f16 = O.faithful16_denoiser_forward(full, to_oracle_cfg(cfg), a["state"], a["action"], a["goal"], a["sigma"])
            line(f"{name} forward [fast] vs 16-bit-faithful oracle", out, f16)
            line(f"{name} (16-bit-faithful oracle vs reference)", f16, a["out"])

print("== samplers (K256 goldens) ==")
cfg, meta, a = load_golden("samplers_K256")
g = {k: v.to(dev) for k, v in a.items() if isinstance(v, torch.Tensor)}
for mode, m in models_for(cfg, golden_weights(cfg, meta)).items():
    use(mode)
    worst = 0.0
    for n in (1, 3, 5):
        for s in ("ddim", "euler", "heun"):
            got = sampling.SAMPLERS[s](m, g["state"], g["x_t"], g["goal"], a[f"sigmas_{n}"]).cpu()
            worst = max(worst, float((got - a[f"{s}_{n}"]).abs().max()))
    print(f"[{mode}] ddim / euler / heun at 1, 3, 5 steps: worst max|err| {worst:.2e}")
