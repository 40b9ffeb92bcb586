"""Data-parallel gradient exchange for BASELINE config 4.

The reference is single-process.  Here every rank computes the local-mean loss and a flat fp32 gradient
on its shard; ONE all-reduce(sum) of that flat buffer followed by a 1/world scale gives exactly the
global-batch mean gradient when shards are equal-sized (the loss is a plain mean, score_wrappers.py:79).

Two transports:
* ``"nccl"``  -- ``beso_comm_*`` / ``beso_allreduce_grads`` of the C ABI: ncclAllReduce on the caller's
  stream over NVLink/NVSwitch, communicator bootstrapped from a unique id shared through torch.distributed.
* ``"torch"`` -- ``torch.distributed.all_reduce`` on whatever backend the process group has (gloo on CPU);
  used by the CPU tests of the host-side logic.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch
import torch.distributed as dist

from . import _lib


def shard_batch(n: int, rank: int, world: int) -> slice:
    """Contiguous equal shards; n must be divisible by world so that mean-of-means == global mean."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return slice(rank * per, (rank + 1) * per)


class FlatGradAllReduce:
    def __init__(self, transport: str = "nccl", device: int | None = None):
        self.transport = transport
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self._comm = None
        if transport == "nccl" and self.world > 1:
            device = torch.cuda.current_device() if device is None else device
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                _lib.check(_lib.lib().beso_comm_unique_id(buf), "beso_comm_unique_id")
            obj = [bytes(buf.raw)]
            dist.broadcast_object_list(obj, src=0)
            handle = C.c_void_p()
            _lib.check(_lib.lib().beso_comm_init(self.rank, self.world, obj[0], device, C.byref(handle)), "beso_comm_init")
            self._comm = handle

    def comm_handle(self):
        """The NCCL communicator of the C ABI when this exchange can run inside ``beso_loss_fwd_bwd_dp`` (overlapped
        with the backward pass), else None."""
        return self._comm if (self.transport == "nccl" and self.world > 1) else None

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        """In place: flat <- sum over ranks(flat) / world."""
        if self.world == 1:
            return flat
        if self.transport == "nccl":
            stream = torch.cuda.current_stream(flat.device).cuda_stream
            _lib.check(_lib.lib().beso_allreduce_grads(self._comm, flat.data_ptr(), flat.numel(), 1.0 / self.world,
                                                      C.c_void_p(stream)), "beso_allreduce_grads")
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.mul_(1.0 / self.world)
        return flat

    def close(self):
        if self._comm is not None:
            _lib.lib().beso_comm_destroy(self._comm)
            self._comm = None


def assign_grads(params: List[torch.nn.Parameter], flat: torch.Tensor):
    """Point every ``p.grad`` at its slice of the (all-reduced) flat buffer, parameters() order."""
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
