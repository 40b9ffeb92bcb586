"""The reference's REAL model shapes (the frozen configs beside the shipped checkpoints: kitchen d = 360 / 6 heads of 60 /
6 layers, block-push d = 240 / 12 heads of 20 / 4 layers) on one B200: every arithmetic mode of the library next to the
reference's algorithm through stock PyTorch eager on the same GPU (the oracle restatement = the same ATen ops).

Workloads per shape: one forward at batch 4096; a 50-step DDIM sample loop at batch 512; the rollout call the reference
actually makes (predict(): batch 1, 10-step DDIM).  CUDA events, warm, best of 5.  Reporting only (lives under tests/
because it imports the oracle).      python tests/perf_checkpoint_shapes.py > gpurun_out/ckpt_shapes.json"""
import json
import sys

import torch

sys.path.insert(0, ".")
from beso_b200 import BLOCKPUSH_CKPT, KITCHEN_CKPT, sampling                    # noqa: E402
from beso_b200.denoiser import build_denoiser                                   # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict              # noqa: E402
from oracle import beso_oracle as O                                             # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = float("inf")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    rows = []
    sig50 = sampling.get_sigmas_exponential(50, 0.005, 1.0)
    sig10 = sampling.get_sigmas_exponential(10, 0.005, 1.0)
    for label, cfg in (("kitchen_ckpt", KITCHEN_CKPT), ("blockpush_ckpt", BLOCKPUSH_CKPT)):
        sd = synthetic_state_dict(cfg, seed=1)
        xs = {b: {k: v.to(DEV) for k, v in synthetic_inputs(cfg, b, seed=2).items()} for b in (1, 512, 4096)}
        fl = cfg.fwd_flops_per_seq()
        for mode in ("fast", "precise", "simt"):
            m = build_denoiser(cfg, DEV, mode=mode, state_dict=sd)
            m.refresh_weights()
            x = xs[4096]
            ms = timed(lambda: m(x["state"], x["action"], x["goal"], x["sigma"]), reps=5 if mode != "simt" else 2, warm=1)
            rows.append(dict(shape=label, mode=mode, workload="fwd b4096", ms=ms, denoise_steps_per_s=4096 / ms * 1e3,
                             tflops=4096 * fl / ms / 1e9))
            x = xs[512]
            ms = timed(lambda: sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig50), reps=3 if mode != "simt" else 1, warm=1)
            rows.append(dict(shape=label, mode=mode, workload="ddim50 b512", ms=ms, denoise_steps_per_s=512 * 50 / ms * 1e3,
                             tflops=512 * 50 * fl / ms / 1e9))
            x = xs[1]
            ms = timed(lambda: sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig10))
            rows.append(dict(shape=label, mode=mode, workload="rollout b1 ddim10", ms=ms, ms_per_evaluation=ms / 10))
        # stock PyTorch eager on the same GPU
        oc = O.OracleCfg(obs_dim=cfg.obs_dim, act_dim=cfg.act_dim, window=cfg.window, goal_len=cfg.goal_len, d=cfg.d,
                         n_layers=cfg.n_layers, n_heads=cfg.n_heads, sigma_data=cfg.sigma_data)
        sdd = {k: v.to(DEV) for k, v in O.as_module_params(sd).items()}
        with torch.no_grad():
            for name, tf32 in (("eager_fp32", False), ("eager_tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                x = xs[4096]
                ms = timed(lambda: O.denoiser_forward(sdd, oc, x["state"], x["action"], x["goal"], x["sigma"]))
                rows.append(dict(shape=label, mode=name, workload="fwd b4096", ms=ms, denoise_steps_per_s=4096 / ms * 1e3))
                x = xs[512]
                ms = timed(lambda: O.sample_ddim(sdd, oc, x["state"], x["noise"], x["goal"], sig50.to(DEV)), reps=3, warm=1)
                rows.append(dict(shape=label, mode=name, workload="ddim50 b512", ms=ms, denoise_steps_per_s=512 * 50 / ms * 1e3))
                x = xs[1]
                ms = timed(lambda: O.sample_ddim(sdd, oc, x["state"], x["noise"], x["goal"], sig10.to(DEV)), reps=3, warm=1)
                rows.append(dict(shape=label, mode=name, workload="rollout b1 ddim10", ms=ms, ms_per_evaluation=ms / 10))
        torch.backends.cuda.matmul.allow_tf32 = False
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
