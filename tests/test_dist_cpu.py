"""CPU, world_size 2, gloo: the host-side data-parallel logic of BASELINE config 4 -- equal shards, one flat
all-reduce, 1/world scale -- reproduces the single-process full-batch gradient.  The per-rank compute is the
oracle (test infrastructure); the exchange is beso_b200.dist."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import to_oracle_cfg
from beso_b200.config import ModelConfig
from beso_b200.dist import FlatGradAllReduce, assign_grads, shard_batch
from beso_b200.synth import synthetic_inputs, synthetic_state_dict

CFG = ModelConfig(obs_dim=12, act_dim=3, window=3, goal_len=1, d=32, n_layers=2, n_heads=2)


def _flat(grads, names):
    return torch.cat([grads[n].reshape(-1) for n in names])


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import beso_oracle as O
    sd = synthetic_state_dict(CFG, 51)
    names = [k for k in sd if not k.endswith("attn.mask")]
    x = synthetic_inputs(CFG, 16, seed=52)
    sl = shard_batch(16, rank, world)
    loss, grads = O.loss_and_grads(sd, to_oracle_cfg(CFG), x["state"][sl], x["clean"][sl], x["goal"][sl],
                                   x["noise"][sl].clone(), x["sigma"][sl])
    flat = _flat(grads, names)
    FlatGradAllReduce("torch")(flat)
    torch.save({"flat": flat, "loss": loss}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_equals_full_batch(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle import beso_oracle as O
    sd = synthetic_state_dict(CFG, 51)
    names = [k for k in sd if not k.endswith("attn.mask")]
    x = synthetic_inputs(CFG, 16, seed=52)
    _, grads = O.loss_and_grads(sd, to_oracle_cfg(CFG), x["state"], x["clean"], x["goal"], x["noise"].clone(), x["sigma"])
    want = _flat(grads, names)
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])                     # replicas stay identical
    torch.testing.assert_close(r0["flat"], want, rtol=1e-4, atol=1e-7)


def test_shard_and_assign():
    assert shard_batch(8192, 3, 8) == slice(3072, 4096)
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
    params = [torch.nn.Parameter(torch.zeros(2, 3)), torch.nn.Parameter(torch.zeros(4))]
    flat = torch.arange(10.0)
    assign_grads(params, flat)
    assert params[0].grad.shape == (2, 3) and params[1].grad.tolist() == [6.0, 7.0, 8.0, 9.0]
    assert FlatGradAllReduce("torch")(flat) is flat               # world size 1: no-op


def _agent_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import beso_b200.training as T
    from beso_b200 import K256
    from beso_b200.agent import BesoAgent
    from beso_b200.denoiser import build_denoiser
    agent = BesoAgent(build_denoiser(K256, "cpu"), device="cpu", window_size=10)
    T.loss_and_flat_grad = lambda *a, **k: (torch.tensor(float(rank)), torch.arange(6.0) * (rank + 1))   # stub: no GPU here
    got = {}
    agent.optimizer = type("O", (), {"step": lambda self, flat_grad=None: got.update(flat=flat_grad.clone())})()
    agent.lr_scheduler = type("S", (), {"step": lambda self: None})()
    agent.ema_helper = type("E", (), {"update": lambda self, p: None})()
    agent.update_ema_every_n_steps, agent.steps = 1, 0
    agent.sigma_sample_density_type, agent.sigma_sample_density_mean, agent.sigma_sample_density_std = "loglogistic", -1.2, 1.2
    agent.enable_data_parallel("torch")
    x = synthetic_inputs(K256, 2, seed=rank)
    loss = agent.train_step({"observation": x["state"], "goal_observation": x["goal"], "action": x["clean"]})
    torch.save({"flat": got["flat"], "loss": loss}, os.path.join(out_dir, f"agent{rank}.pt"))
    dist.destroy_process_group()


def test_agent_train_step_averages_gradients_over_ranks(tmp_path):
    """BesoAgent.enable_data_parallel: the optimiser of every rank sees the mean of the ranks' flat gradients (the
    fused loss + backward is stubbed; the exchange is the real torch.distributed all-reduce over gloo)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_agent_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(os.path.join(tmp_path, "agent0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "agent1.pt"))
    want = torch.arange(6.0) * 1.5                                  # mean of 1x and 2x
    assert torch.equal(r0["flat"], want) and torch.equal(r1["flat"], want)
    assert (r0["loss"], r1["loss"]) == (0.0, 1.0)                   # the loss each rank reports is its local one
