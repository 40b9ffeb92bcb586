"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck): one forward and one 2-step sample loop of the
tensor-core kernel in the selected arithmetic mode, or one small training step (tcgen05 GEMMs + backward kernels).

    compute-sanitizer --tool racecheck python tools/sanitize_target.py fast|precise|simt|train|wide_fast|wide_precise
    BESO_FAST_CG=2 / BESO_FAST_MC=2 select the cluster variants of the fp16 kernel; BESO_PREC_LAYOUT=stacked|p128 the tile
    layout of the precise mode (p128 = 128-row tiles, three MMAs per product); wide_* = the 384-column geometry (d = 360).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import K256                                     # noqa: E402
from beso_b200.cfg import ClassifierFreeSampleModel            # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.sampling import get_sigmas_exponential, sample_ddim, sample_heun  # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fast"
dev = torch.device("cuda:0")
if what.startswith("wide_"):                                   # the reference's kitchen checkpoint width, two layers
    from beso_b200.config import ModelConfig
    K256 = ModelConfig(obs_dim=30, act_dim=9, window=4, goal_len=2, d=360, n_layers=2, n_heads=6)
    what = what[5:]
sd = synthetic_state_dict(K256, 1)
if what == "train":
    from beso_b200.training import loss_and_flat_grad
    m = build_denoiser(K256, dev, mode="precise", state_dict=sd, attn_pdrop=0.3)
    m.train()
    g = {k: v.to(dev) for k, v in synthetic_inputs(K256, 24, seed=2).items()}
    from beso_b200.training import draw_dropout_masks
    for math in ("fp32", "bf16x2", "bf16"):
        m.train_math = math
        masks = draw_dropout_masks(m.inner_model, 24, K256.window, dev)
        loss, flat = loss_and_flat_grad(m, g["state"], g["clean"], g["goal"], g["noise"], g["sigma"], dropout_masks=masks)
        print(math, float(loss), float(flat.abs().sum()))
else:
    m = build_denoiser(K256, dev, mode=what, state_dict=sd)
    g = {k: v.to(dev) for k, v in synthetic_inputs(K256, 13, seed=2).items()}      # ragged: 3 tiles, the last not full
    out = m(g["state"], g["action"], g["goal"], g["sigma"])
    sig = get_sigmas_exponential(2, 0.005, 1.0)
    x = sample_ddim(m, g["state"], g["noise"], g["goal"], sig)
    y = sample_heun(ClassifierFreeSampleModel(m, 2.0), g["state"], g["noise"], g["goal"], sig)
    print(what, float(out.abs().sum()), float(x.abs().sum()), float(y.abs().sum()))
torch.cuda.synchronize()
print("ok")
