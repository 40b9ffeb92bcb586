"""Device-time numbers for every BASELINE.json config on one GPU (CUDA events, warm L2, best of 5).

    python tools/configs_bench.py > gpurun_out/configs.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import B256, K256, T16                                  # noqa: E402
from beso_b200.cfg import ClassifierFreeSampleModel                    # noqa: E402
from beso_b200.denoiser import build_denoiser                          # noqa: E402
from beso_b200.optim import ExponentialMovingAverage, FusedAdamW       # noqa: E402
from beso_b200 import sampling                                         # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict     # noqa: E402
from beso_b200.training import loss_and_flat_grad                      # noqa: E402

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0


def best_ms(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return min(out)


def main():
    res = {}
    # cfg1: single denoise step, batch 64
    for mode in ("precise", "fast"):
        m = build_denoiser(K256, dev, mode=mode, state_dict=synthetic_state_dict(K256, 1))
        x = {k: v.to(dev) for k, v in synthetic_inputs(K256, 64, seed=2).items()}
        ms = best_ms(lambda: m(x["state"], x["action"], x["goal"], x["sigma"]))
        res[f"cfg1_fwd_b64_{mode}"] = {"ms": ms, "denoise_steps_per_s": 64 / ms * 1e3}
    for mode in ("fast", "precise"):
        m = build_denoiser(K256, dev, mode=mode, state_dict=synthetic_state_dict(K256, 1))
        # batch 1 rollout step: 10-step DDIM / euler_ancestral at B = 1 (predict())
        x = {k: v.to(dev) for k, v in synthetic_inputs(K256, 1, seed=2).items()}
        sig = sampling.get_sigmas_exponential(10, 0.005, 1.0)
        res[f"rollout_b1_ddim10_{mode}"] = {"ms": best_ms(lambda: sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig))}
        res[f"rollout_b1_euler_ancestral10_{mode}"] = {"ms": best_ms(lambda: sampling.sample_euler_ancestral(m, x["state"], x["noise"], x["goal"], sig))}
        # cfg2: 50-step DDIM, batch 512
        x = {k: v.to(dev) for k, v in synthetic_inputs(K256, 512, seed=2).items()}
        sig50 = sampling.get_sigmas_exponential(50, 0.005, 1.0)
        ms = best_ms(lambda: sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig50))
        fl = 512 * 50 * K256.fwd_flops_per_seq()
        res[f"cfg2_ddim50_b512_{mode}"] = {"ms": ms, "denoise_steps_per_s": 512 * 50 / ms * 1e3, "tflops": fl / ms / 1e9, "frac_sustained": fl / ms / 1e9 / PEAK}
        # cfg5: CFG (lambda 2), 10-step Heun, batch 2048: 19 sampler evaluations x 2 branches
        x = {k: v.to(dev) for k, v in synthetic_inputs(K256, 2048, seed=2).items()}
        w = ClassifierFreeSampleModel(m, cond_lambda=2.0)
        sig10 = sampling.get_sigmas_exponential(10, 0.005, 1.0)
        ms = best_ms(lambda: sampling.sample_heun(w, x["state"], x["noise"], x["goal"], sig10))
        evals = 2048 * 19 * 2
        fl = evals * K256.fwd_flops_per_seq()
        res[f"cfg5_cfg_heun10_b2048_{mode}"] = {"ms": ms, "model_evals_per_s": evals / ms * 1e3, "sampler_steps_per_s": 2048 * 10 / ms * 1e3,
                                               "tflops": fl / ms / 1e9, "frac_sustained": fl / ms / 1e9 / PEAK}
        # north star: one forward, batch 4096
        for name, c in ((f"ns_fwd_T16_b4096_{mode}", T16), (f"ns_fwd_K256_b4096_{mode}", K256)):
            mm = build_denoiser(c, dev, mode=mode, state_dict=synthetic_state_dict(c, 3))
            xi = {k: v.to(dev) for k, v in synthetic_inputs(c, 4096, seed=4).items()}
            ms = best_ms(lambda: mm(xi["state"], xi["action"], xi["goal"], xi["sigma"]))
            fl = 4096 * c.fwd_flops_per_seq()
            res[name] = {"ms": ms, "denoise_steps_per_s": 4096 / ms * 1e3, "tflops": fl / ms / 1e9, "frac_sustained": fl / ms / 1e9 / PEAK}
    # cfg3: block-push training step, batch 4096 (loss + backward + fused AdamW/EMA step)
    for math in ("fp32", "bf16x2", "bf16"):
        mt = build_denoiser(B256, dev, mode="precise", state_dict=synthetic_state_dict(B256, 41))
        mt.train(); mt.train_math = math
        g = {k: v.to(dev) for k, v in synthetic_inputs(B256, 4096, seed=42, sigma_min=0.05).items()}
        params = list(mt.inner_model.parameters())
        opt = FusedAdamW(params, lr=1e-4)
        ema = ExponentialMovingAverage(params, 0.999)
        opt.attach_ema(ema)

        def step():
            _, flat = loss_and_flat_grad(mt, g["state"], g["clean"], g["goal"], g["noise"], g["sigma"])
            opt.step(flat_grad=flat)
            ema.update(params)
        ms = best_ms(step, reps=3, warm=1)
        res[f"cfg3_train_step_b4096_{math}"] = {"ms": ms, "samples_per_s": 4096 / ms * 1e3, "tflops": 3 * 4096 * B256.fwd_flops_per_seq() / ms / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
