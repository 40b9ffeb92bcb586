"""GPU-resident windowed trajectory dataset.

Mirror of ``TrajectorySlicerDataset`` (beso/envs/dataloaders/trajectory_loader.py:79-197) for the training path: the
padded trajectories live in HBM once, and a whole training batch -- the reference's per-sample ``__getitem__`` slicing,
the DataLoader's dict collation and the host-to-device copy in ``BesoAgent.train_step`` -- is ONE launch of
``beso_window_gather`` (beso_b200/csrc/dataset.cu).  Slice enumeration order, the future-window rules and the order of
``np.random.randint`` draws follow the reference, so that with the same numpy seed and the same index order the
batches are identical to the reference's.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


class DeviceWindowDataset:
    """``observations`` (N, t_max, obs_dim) and ``actions`` (N, t_max, act_dim) are the padded trajectories, ``seq_lengths``
    (N,) their valid lengths (``TrajectoryDataset.get_seq_length``).  Item ``idx`` is the window ``slices[idx] = (i, start,
    start + window)``; ``get_batch`` returns the dict the reference's DataLoader yields (``observation``, ``action`` and,
    when ``future_conditional``, ``goal_observation``), already on the device."""

    def __init__(self, observations, actions, seq_lengths: Sequence[int], window: int, future_conditional: bool = False,
                 min_future_sep: int = 0, future_seq_len: Optional[int] = None, only_sample_tail: bool = False,
                 only_sample_seq_end: bool = False, device="cuda", scaler=None):
        if future_conditional and future_seq_len is None:
            raise AssertionError("must specify a future_seq_len")  # trajectory_loader.py:116-117
        self.device = torch.device(device)
        self.observations = torch.as_tensor(observations, dtype=torch.float32).to(self.device).contiguous()
        self.actions = torch.as_tensor(actions, dtype=torch.float32).to(self.device).contiguous()
        if self.observations.dim() != 3 or self.actions.dim() != 3 or self.observations.shape[:2] != self.actions.shape[:2]:
            raise ValueError("observations / actions must be (N, t_max, dim) with the same N and t_max")
        self.seq_lengths = np.asarray(seq_lengths, dtype=np.int64)
        if self.seq_lengths.shape != (self.observations.shape[0],) or (self.seq_lengths > self.observations.shape[1]).any():
            raise ValueError("seq_lengths must hold one length <= t_max per trajectory")
        self.window = int(window)
        self.future_conditional = bool(future_conditional)
        self.min_future_sep = int(min_future_sep)
        self.future_seq_len = None if future_seq_len is None else int(future_seq_len)
        self.only_sample_tail = bool(only_sample_tail)
        self.only_sample_seq_end = bool(only_sample_seq_end)
        self.scaler = scaler
        # slices in the reference's order (trajectory_loader.py:130-139): trajectory-major, then start; short ones skipped
        traj, start = [], []
        for i, T in enumerate(self.seq_lengths.tolist()):
            if T - self.window >= 0:
                n = T - self.window + 1
                traj.append(np.full(n, i, dtype=np.int32))
                start.append(np.arange(n, dtype=np.int32))
        self.slice_traj = np.concatenate(traj) if traj else np.zeros(0, np.int32)
        self.slice_start = np.concatenate(start) if start else np.zeros(0, np.int32)

    def __len__(self) -> int:
        return int(self.slice_traj.shape[0])

    def get_seq_length(self, idx: int) -> int:  # trajectory_loader.py:146-150
        return self.future_seq_len + self.window if self.future_conditional else self.window

    def goal_starts(self, indices, rng=None) -> np.ndarray:
        """Start of the future-observation window of each item, in ``indices`` order, -1 for the zeros placeholder
        (trajectory_loader.py:170-184).  ``rng`` is ``np.random`` (the reference's global generator) or a RandomState;
        one ``randint`` is drawn per item that samples, in order, exactly as iterating the reference dataset does."""
        rng = np.random if rng is None else rng
        indices = np.asarray(indices, dtype=np.int64).reshape(-1)
        G, t_max = self.future_seq_len, int(self.observations.shape[1])
        traj = self.slice_traj[indices].astype(np.int64)
        end = self.slice_start[indices].astype(np.int64) + self.window
        lo, hi = end + self.min_future_sep, self.seq_lengths[traj] - G
        ok = lo < hi
        out = np.full(indices.shape[0], -1, dtype=np.int32)
        if self.only_sample_tail:
            out[ok] = t_max - G                              # the reference slices the PADDED trajectory's last G frames
        elif self.only_sample_seq_end:
            out[ok] = end[ok]
        elif ok.any():
            # one legacy-generator call with array bounds draws exactly what successive scalar randint(lo, hi) calls
            # draw, element by element, and leaves the generator in the same state (tests/test_dataset.py)
            out[ok] = rng.randint(lo[ok], hi[ok])
        return out

    def attach_scaler(self, scaler) -> None:
        """Batches come out already scaled (``scaler.scale_input`` on observations and goals, ``scale_output`` on
        actions, fused into the gather) and carry ``"scaled": True`` so that ``BesoAgent.train_step`` skips its own
        scaling."""
        self.scaler = scaler

    def get_batch(self, indices, goal_start=None, rng=None, stream=None) -> dict:
        """One launch: the collated batch of ``indices`` (host ints).  ``goal_start`` overrides ``goal_starts``."""
        indices = np.asarray(indices, dtype=np.int64).reshape(-1)
        if indices.size == 0:
            raise ValueError("empty batch")
        if (indices < 0).any() or (indices >= len(self)).any():
            raise IndexError("window index out of range")
        B, W = int(indices.size), self.window
        obs_dim, act_dim = int(self.observations.shape[2]), int(self.actions.shape[2])
        meta = [self.slice_traj[indices], self.slice_start[indices]]
        G = 0
        if self.future_conditional:
            G = self.future_seq_len
            gs = self.goal_starts(indices, rng) if goal_start is None else np.asarray(goal_start, dtype=np.int32).reshape(-1)
            if gs.shape[0] != B or (gs.astype(np.int64) + G > self.observations.shape[1]).any():
                raise ValueError("goal_start must hold one start per item with start + future_seq_len <= t_max")
            meta.append(gs)
        meta_dev = torch.from_numpy(np.stack(meta).astype(np.int32)).to(self.device, non_blocking=True)
        out = {"observation": torch.empty(B, W, obs_dim, device=self.device),
               "action": torch.empty(B, W, act_dim, device=self.device)}
        goal_ptr, gs_ptr = None, None
        obs_tab, act_tab = (None, None) if self.scaler is None else self.scaler.gather_tables()
        if self.scaler is not None:
            out["scaled"] = True
        if self.future_conditional:
            out["goal_observation"] = torch.empty(B, G, obs_dim, device=self.device)
            goal_ptr, gs_ptr = out["goal_observation"].data_ptr(), meta_dev[2].data_ptr()
        s = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _lib.check(_lib.lib().beso_window_gather(
            self.observations.data_ptr(), self.actions.data_ptr(), int(self.observations.shape[0]),
            int(self.observations.shape[1]), obs_dim, act_dim, meta_dev[0].data_ptr(), meta_dev[1].data_ptr(), gs_ptr, W, G,
            out["observation"].data_ptr(), out["action"].data_ptr(), goal_ptr,
            None if obs_tab is None else obs_tab.data_ptr(), None if act_tab is None else act_tab.data_ptr(), B,
            C.c_void_p(s)), "beso_window_gather")
        return out

    def epoch_order(self, shuffle: bool = True, generator=None, rank: int = 0, world_size: int = 1) -> np.ndarray:
        """Window indices of one epoch for this rank.  One rank: ``torch.randperm`` (what the DataLoader's RandomSampler
        draws) or in order.  Data-parallel training: every rank draws the SAME permutation (same
        generator seed), pads it by wrapping to a multiple of ``world_size`` and takes ``order[rank::world_size]`` --
        the rule of ``torch.utils.data.DistributedSampler`` -- so shards are disjoint, equal-sized, and the mean of the
        per-rank gradients is the global-batch gradient."""
        n = len(self)
        if shuffle and world_size > 1 and generator is None:
            # the global generator differs between ranks: their "same" permutation would not be the same and the
            # shards would overlap silently
            raise ValueError("data-parallel shuffling needs an explicit generator seeded identically on every rank")
        order = torch.randperm(n, generator=generator).numpy() if shuffle else np.arange(n)
        if world_size > 1:
            if not 0 <= rank < world_size:
                raise ValueError("rank must be in [0, world_size)")
            total = -(-n // world_size) * world_size
            if total > n:
                order = np.concatenate([order, np.resize(order, total - n)])
            order = order[rank:total:world_size]
        return order

    def batches(self, batch_size: int, shuffle: bool = True, drop_last: bool = False, generator=None, rng=None,
                rank: int = 0, world_size: int = 1):
        """Epoch iterator over ``epoch_order``; with ``world_size`` > 1 each rank gathers only its own shard."""
        order = self.epoch_order(shuffle, generator, rank, world_size)
        for lo in range(0, order.shape[0], batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.shape[0] < batch_size:
                return
            yield self.get_batch(idx, rng=rng)
