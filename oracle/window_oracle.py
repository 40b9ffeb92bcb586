"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's windowed trajectory dataset.

Follows ``TrajectorySlicerDataset`` (beso/envs/dataloaders/trajectory_loader.py:79-197) over padded arrays
``obs (N, t_max, obs_dim)``, ``act (N, t_max, act_dim)`` and valid lengths ``lens (N,)``.  Parity status: PINNED --
tests/golden/windows.npz holds batches produced by the unmodified reference class (oracle/make_golden.py windows), and
tests/test_dataset.py also runs the reference class live when /root/reference is present.

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def slices(lens, window):
    """trajectory_loader.py:130-139: (i, start, end) for every start in range(T - window + 1), trajectory-major."""
    out = []
    for i, T in enumerate(np.asarray(lens).tolist()):
        if T - window >= 0:
            out += [(i, s, s + window) for s in range(T - window + 1)]
    return out


def item(obs, act, lens, sl, idx, window, future_conditional=False, min_future_sep=0, future_seq_len=None,
         only_sample_tail=False, only_sample_seq_end=False, rng=np.random):
    """trajectory_loader.py:160-197 (``__getitem__``), the transform hook left out (the reference returns the dict)."""
    i, start, end = sl[idx]
    out = {"observation": obs[i, start:end], "action": act[i, start:end]}
    if future_conditional:
        lo, hi = end + min_future_sep, int(lens[i]) - future_seq_len
        if lo < hi:
            if only_sample_tail:
                fut = obs[i, -future_seq_len:]           # last frames of the padded trajectory, as the reference does
            elif only_sample_seq_end:
                fut = obs[i, end:end + future_seq_len]
            else:
                s = rng.randint(lo, hi)
                fut = obs[i, s:s + future_seq_len]
        else:
            fut = np.zeros((future_seq_len, obs.shape[2]), dtype=obs.dtype)
        out["goal_observation"] = fut
    return out


def batch(obs, act, lens, indices, window, **kw):
    """The default DataLoader collation of ``item`` over ``indices`` (a stack per key, in index order)."""
    sl = slices(lens, window)
    items = [item(obs, act, lens, sl, int(i), window, **kw) for i in indices]
    return {k: np.stack([it[k] for it in items]) for k in items[0]}
