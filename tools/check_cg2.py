"""Parity of the cluster variants of the FAST path against the PRECISE kernel: run with BESO_FAST_CG=2
(cta_group::2 MMAs over a CTA pair) or BESO_FAST_MC=2 (independent CTAs sharing the weight stream by TMA
multicast); odd and even tile counts, so that a cluster with a dummy tile is covered."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beso_b200 import K256, T16                                  # noqa: E402
from beso_b200.denoiser import build_denoiser                    # noqa: E402
from beso_b200.sampling import get_sigmas_exponential, sample_heun  # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402

assert os.environ.get("BESO_FAST_CG") == "2" or os.environ.get("BESO_FAST_MC") == "2"
dev = torch.device("cuda:0")
for cfg, B in ((K256, 37), (T16, 64), (K256, 1500)):
    sd = synthetic_state_dict(cfg, 31)
    x = {k: v.to(dev) for k, v in synthetic_inputs(cfg, B, seed=32).items()}
    fast = build_denoiser(cfg, dev, mode="fast", state_dict=sd)
    prec = build_denoiser(cfg, dev, mode="precise", state_dict=sd)
    a = fast(x["state"], x["action"], x["goal"], x["sigma"])
    b = prec(x["state"], x["action"], x["goal"], x["sigma"])
    torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)
    sig = get_sigmas_exponential(3, 0.005, 1.0)
    a = sample_heun(fast, x["state"], x["noise"], x["goal"], sig)
    b = sample_heun(prec, x["state"], x["noise"], x["goal"], sig)
    torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)
print("cg2 ok")
