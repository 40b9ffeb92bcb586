"""Times the window-gather launch (beso_b200/csrc/dataset.cu) at training batch sizes with CUDA events and reports the
achieved HBM rate (algorithmic bytes = one read + one write of every output float).  Also times the host-side slicing
the reference does for the same batch (numpy restatement, per-sample indexing + stacking + one H2D copy)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from beso_b200 import scaler as S                      # noqa: E402
from beso_b200.dataset import DeviceWindowDataset      # noqa: E402
from oracle import window_oracle as WO                 # noqa: E402  (checker / CPU baseline only)


def main():
    rs = np.random.RandomState(0)
    N, t_max, obs_dim, act_dim = 566, 409, 60, 9          # relay-kitchen sized: 566 demonstrations, up to 409 frames
    lens = rs.randint(150, t_max + 1, size=N)
    obs, act = rs.randn(N, t_max, obs_dim).astype(np.float32), rs.randn(N, t_max, act_dim).astype(np.float32)
    kw = dict(window=5, future_conditional=True, min_future_sep=10, future_seq_len=2)
    sc = S.Scaler(obs, act, True, "cuda")
    ds = DeviceWindowDataset(obs, act, lens, device="cuda", scaler=sc, **kw)
    out = {"windows": len(ds), "resident_mb": (obs.nbytes + act.nbytes) / 2**20, "rows": []}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for B in (1024, 16384, 262144):
        idx = rs.randint(0, len(ds), size=B)
        gs = ds.goal_starts(idx, np.random.RandomState(1))
        ds.get_batch(idx, goal_start=gs)
        times = []
        for _ in range(10):
            flush.fill_(1)
            t0 = time.perf_counter()
            b = ds.get_batch(idx, goal_start=gs)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        # kernel alone: metadata already on the device is not separable through the public call, so time launches back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            b = ds.get_batch(idx, goal_start=gs)
        e1.record()
        torch.cuda.synchronize()
        floats = sum(v.numel() for k, v in b.items() if k != "scaled")
        row = {"batch": B, "call_ms_median": float(np.median(times) * 1e3), "pipelined_ms_per_call": e0.elapsed_time(e1) / 20,
               "algorithmic_bytes": floats * 8}
        row["pipelined_GBps"] = row["algorithmic_bytes"] / (row["pipelined_ms_per_call"] * 1e-3) / 1e9
        if B <= 16384:
            t0 = time.perf_counter()
            ref = WO.batch(obs, act, lens, idx, kw["window"], future_conditional=True, min_future_sep=10, future_seq_len=2,
                           rng=np.random.RandomState(1))
            dev = {k: torch.from_numpy(v).cuda() for k, v in ref.items()}
            dev = {"observation": sc.scale_input(dev["observation"]), "goal_observation": sc.scale_input(dev["goal_observation"]),
                   "action": sc.scale_output(dev["action"])}
            torch.cuda.synchronize()
            row["host_slicing_ms"] = (time.perf_counter() - t0) * 1e3
            b2 = ds.get_batch(idx, rng=np.random.RandomState(1))
            row["bit_exact_vs_host"] = all(torch.equal(b2[k], dev[k]) for k in dev)
        out["rows"].append(row)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
