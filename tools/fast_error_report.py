"""Measured error of the FAST mode against the reference golden fixtures (forward, samplers).  python tools/fast_error_report.py"""
import glob
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import golden_weights, load_golden                 # noqa: E402
from beso_b200 import sampling                                   # noqa: E402
from beso_b200.denoiser import build_denoiser                    # noqa: E402

dev = torch.device("cuda:0")
worst = 0.0
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fwd_*.npz"))):
    name = os.path.basename(path)[:-4]
    cfg, meta, a = load_golden(name)
    try:
        m = build_denoiser(cfg, dev, mode="fast", state_dict=golden_weights(cfg, meta))
        out = m(a["state"].to(dev), a["action"].to(dev), a["goal"].to(dev), a["sigma"].to(dev)).cpu()
    except Exception as e:                                       # shapes the fast kernel does not support
        print(f"{name}: skipped ({type(e).__name__})")
        continue
    err = (out - a["out"]).abs()
    need = (err / (1e-3 * a["out"].abs() + 1e-5)).max()
    print(f"{name}: max|err| {float(err.max()):.2e} mean {float(err.mean()):.2e} max|ref| {float(a['out'].abs().max()):.2f} "
          f"inside 1e-3/1e-5: {float((err <= 1e-5 + 1e-3 * a['out'].abs()).float().mean()):.2f} worst err/(tol) {float(need):.1f}")
    worst = max(worst, float(err.max()))
cfg, meta, a = load_golden("samplers_K256")
m = build_denoiser(cfg, dev, mode="fast", state_dict=golden_weights(cfg, meta))
g = {k: v.to(dev) for k, v in a.items() if isinstance(v, torch.Tensor)}
for n in (1, 3, 5):
    for s in ("ddim", "euler", "heun"):
        got = sampling.SAMPLERS[s](m, g["state"], g["x_t"], g["goal"], a[f"sigmas_{n}"]).cpu()
        err = (got - a[f"{s}_{n}"]).abs()
        worst = max(worst, float(err.max()))
        print(f"{s}_{n}: max|err| {float(err.max()):.2e} mean {float(err.mean()):.2e}")
print(f"worst max|err| {worst:.2e}")
