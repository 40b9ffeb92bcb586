"""One cfg3 training step (fwd + bwd, B = 4096) for kernel-level profiling.  python tools/train_prof.py [fp32|bf16]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import B256                                    # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402
from beso_b200.training import loss_and_flat_grad             # noqa: E402

math = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dev = torch.device("cuda:0")
m = build_denoiser(B256, dev, mode="precise", state_dict=synthetic_state_dict(B256, 41))
m.train()
m.train_math = math
g = {k: v.to(dev) for k, v in synthetic_inputs(B256, 4096, seed=42, sigma_min=0.05).items()}
for _ in range(2):
    loss, flat = loss_and_flat_grad(m, g["state"], g["clean"], g["goal"], g["noise"], g["sigma"])
torch.cuda.synchronize()
print(float(loss))
