"""GPU diagnostics for the FAST (tcgen05) kernel: compares the residual stream of tile 0 at every
LayerNorm with the oracle's, then the outputs.  Run on the GPU box: python tests/debug_fast_vs_oracle.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import K256, T16, B256, _lib                    # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402
from oracle import beso_oracle as O                           # noqa: E402


def main():
    dev = torch.device("cuda:0")
    for name, cfg, B in (("K256", K256, 7), ("T16", T16, 20), ("B256", B256, 5)):
        sd = synthetic_state_dict(cfg, seed=1)
        x = synthetic_inputs(cfg, B, seed=11)
        oc = O.OracleCfg(obs_dim=cfg.obs_dim, act_dim=cfg.act_dim, window=cfg.window, goal_len=cfg.goal_len,
                         d=cfg.d, n_layers=cfg.n_layers, n_heads=cfg.n_heads, sigma_data=cfg.sigma_data)
        trace = []
        with torch.no_grad():
            c_in = O.get_scalings(x["sigma"], cfg.sigma_data)[2].view(-1, 1, 1)
            O.gpt_forward(sd, oc, x["state"], x["action"] * c_in, x["goal"], x["sigma"], trace=trace)
            want = O.denoiser_forward(sd, oc, x["state"], x["action"], x["goal"], x["sigma"])
        m = build_denoiser(cfg, dev, mode="fast", state_dict=sd)
        g = {k: v.to(dev) for k, v in x.items()}
        n_slots = 2 * cfg.n_layers + 1
        tr = torch.full((n_slots, 128, 256), float("nan"), device=dev)
        _lib.lib().beso_debug_set_trace(C.c_void_p(tr.data_ptr()))
        got = m(g["state"], g["action"], g["goal"], g["sigma"])
        torch.cuda.synchronize()
        _lib.lib().beso_debug_set_trace(None)
        T = cfg.n_tokens()
        S = min(128 // T, B)
        tr = tr.cpu()
        print(f"== {name}: T={T} seqs/tile={128 // T}")
        for i in range(n_slots):
            ref = trace[i][:S].reshape(S * T, 256)
            mine = tr[i, :S * T]
            err = (mine - ref).abs()
            print(f"  slot {i}: max|err|={err.max():.3e} mean|err|={err.mean():.3e} ref_rms={ref.pow(2).mean().sqrt():.3e} "
                  f"nan={int(torch.isnan(mine).sum())}")
        err = (got.cpu() - want).abs()
        print(f"  output: max|err|={err.max():.3e} mean|err|={err.mean():.3e} "
              f"within(1e-3,1e-5)={(err <= 1e-5 + 1e-3 * want.abs()).float().mean():.3f} "
              f"within(2e-2)={(err <= 2e-2 + 2e-2 * want.abs()).float().mean():.3f}")


if __name__ == "__main__":
    main()
