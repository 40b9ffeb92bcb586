"""Runs a few forwards (for ncu captures / quick timing); BESO_RUN_MODE=precise selects the fp32 kernel.

    python tools/run_fwd.py [T16|K256|KITCHEN|PUSH] [B] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import BLOCKPUSH_CKPT, K256, KITCHEN_CKPT, T16    # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "T16"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    cfg = {"K256": K256, "T16": T16, "KITCHEN": KITCHEN_CKPT, "PUSH": BLOCKPUSH_CKPT}[name]
    dev = torch.device("cuda:0")
    m = build_denoiser(cfg, dev, mode=os.environ.get("BESO_RUN_MODE", "fast"), state_dict=synthetic_state_dict(cfg, 1))
    x = {k: v.to(dev) for k, v in synthetic_inputs(cfg, B, seed=2).items()}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    with torch.no_grad():
        m(x["state"], x["action"], x["goal"], x["sigma"])
        ev[0].record()
        for i in range(reps):
            m(x["state"], x["action"], x["goal"], x["sigma"])
            ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(reps)]
    fl = cfg.fwd_flops_per_seq() * B
    print(f"{name} B={B}: ms per forward {['%.4f' % t for t in ms]}  best {min(ms):.4f} ms = {fl / min(ms) / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    main()
