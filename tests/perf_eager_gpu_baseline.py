"""Like-for-like GPU baseline: the reference's algorithm run through stock PyTorch eager ON THE B200
(the oracle restatement = the same ATen ops in the same order; /root/reference itself cannot travel to the GPU box),
fp32 (TF32 off, torch's default), fp32 with TF32 matmuls, and autocast(bf16), on the bench workload (cfg2: 50-step DDIM,
batch 512) and on the single forward at batch 4096.  Timed with CUDA events after warm-up.  Reporting only: nothing
in the product path uses it.   Run on the GPU box: python tests/perf_eager_gpu_baseline.py > gpurun_out/eager_gpu.json"""
import json
import sys

import torch

sys.path.insert(0, ".")
from beso_b200 import K256, T16                                         # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict      # noqa: E402
from oracle import beso_oracle as O                                     # noqa: E402

N_STEPS, SIGMA_MIN, SIGMA_MAX = 50, 0.005, 1.0
DEV = "cuda" if torch.cuda.is_available() else "cpu"     # cpu: dry run of the script logic only
B_LOOP, B_FWD = (512, 4096) if DEV == "cuda" else (4, 8)


def setup(cfg, batch):
    oc = O.OracleCfg(obs_dim=cfg.obs_dim, act_dim=cfg.act_dim, window=cfg.window, goal_len=cfg.goal_len, d=cfg.d,
                     n_layers=cfg.n_layers, n_heads=cfg.n_heads, sigma_data=cfg.sigma_data)
    sd = {k: v.to(DEV) for k, v in O.as_module_params(synthetic_state_dict(cfg, seed=1)).items()}
    x = {k: v.to(DEV) for k, v in synthetic_inputs(cfg, batch, seed=2).items()}
    return oc, sd, x


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    if DEV != "cuda":
        return 1.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0) if DEV == "cuda" else "none", "rows": []}
    sig = O.get_sigmas_exponential(N_STEPS, SIGMA_MIN, SIGMA_MAX).to(DEV)
    modes = {"fp32": (False, None), "fp32+tf32": (True, None), "autocast_bf16": (False, torch.bfloat16)}
    with torch.no_grad():
        for name, (tf32, ac) in modes.items():
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            ctx = (lambda: torch.autocast(DEV, dtype=ac)) if ac is not None else (lambda: torch.autocast(DEV, enabled=False))
            oc, sd, x = setup(K256, B_LOOP)

            def loop():
                with ctx():
                    O.sample_ddim(sd, oc, x["state"], x["noise"] * SIGMA_MAX, x["goal"], sig)
            ms = timed(loop)
            out["rows"].append({"workload": "cfg2 ddim50 b512 K256", "mode": name, "ms": ms,
                                "denoise_steps_per_s": B_LOOP * N_STEPS / (ms * 1e-3)})
            for cfg, label in ((T16, "fwd T16 b4096"), (K256, "fwd K256 b4096")):
                oc, sd, x = setup(cfg, B_FWD)
                s = torch.full((B_FWD,), 0.3, device=DEV)

                def fwd():
                    with ctx():
                        O.denoiser_forward(sd, oc, x["state"], x["noise"], x["goal"], s)
                ms = timed(fwd, reps=10)
                out["rows"].append({"workload": label, "mode": name, "ms": ms, "denoise_steps_per_s": B_FWD / (ms * 1e-3)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
