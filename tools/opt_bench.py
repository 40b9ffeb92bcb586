"""Optimiser + EMA step time at the K256 parameter set: fused launch vs torch.optim.AdamW + the reference's
per-parameter EMA loop.  python tools/opt_bench.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import K256                                   # noqa: E402
from beso_b200.optim import ExponentialMovingAverage, FusedAdamW   # noqa: E402
from beso_b200.synth import synthetic_state_dict              # noqa: E402


def main():
    dev = torch.device("cuda:0")
    sd = synthetic_state_dict(K256, seed=1)
    names = [n for n, _ in K256.param_shapes()]
    make = lambda: [torch.nn.Parameter(sd[n].to(dev).clone()) for n in names]      # noqa: E731
    n = sum(p.numel() for p in make())
    flat = torch.randn(n, device=dev) * 1e-2

    def timed(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    ours = make()
    opt = FusedAdamW(ours, lr=1e-4)
    ema = ExponentialMovingAverage(ours, 0.999)
    opt.attach_ema(ema)

    def fused():
        opt.step(flat_grad=flat)
        ema.update(ours)

    for label, kw in (("foreach", dict(foreach=True)), ("single-tensor", dict(foreach=False, fused=False)), ("torch fused", dict(fused=True))):
        ref = make()
        off = 0
        for q in ref:
            q.grad = flat[off:off + q.numel()].view_as(q).clone()
            off += q.numel()
        o = torch.optim.AdamW(ref, lr=1e-4, **kw)
        shadow = [p.detach().clone() for p in ref]

        def torch_step():
            o.step()
            with torch.no_grad():
                for s, p in zip(shadow, ref):            # ema.py:51-53
                    s.sub_(0.001 * (s - p))
        print(f"torch AdamW ({label}) + per-parameter EMA loop: {timed(torch_step):.3f} ms per step")
    ms = timed(fused)
    print(f"fused beso_opt_step (AdamW + EMA, 1 launch): {ms:.3f} ms per step; {n} parameters, "
          f"{9 * 4 * n / (ms * 1e-3) / 1e9:.0f} GB/s of the 9 x 4 B per element")


if __name__ == "__main__":
    main()
