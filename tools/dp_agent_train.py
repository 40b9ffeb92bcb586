"""Multi-GPU check of the whole data-parallel training slice (BASELINE config 4), one rank per GPU under torchrun:

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tools/dp_agent_train.py

Every rank keeps the same trajectories resident (DeviceWindowDataset, scaler fused into the gather), takes its
DistributedSampler-style shard of each epoch's permutation, and runs BesoAgent.train_step with the flat-gradient
all-reduce (ncclAllReduce through the C ABI) + fused AdamW / EMA.  Checks: the replicas' parameters and EMA weights stay
bit-identical, the loss goes down; prints samples/s over all ranks (device time, max over ranks)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beso_b200 import K256                                          # noqa: E402
from beso_b200 import scaler as S                                   # noqa: E402
from beso_b200.agent import BesoAgent                               # noqa: E402
from beso_b200.dataset import DeviceWindowDataset                   # noqa: E402
from beso_b200.denoiser import build_denoiser                       # noqa: E402
from beso_b200.synth import synthetic_state_dict                    # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg, per_rank = K256, 1024
    rs = np.random.RandomState(0)                                     # the same synthetic demonstrations on every rank
    n_traj, t_max = 128, 300
    lens = rs.randint(100, t_max + 1, size=n_traj)
    obs = (rs.randn(n_traj, t_max, cfg.obs_dim) * 2 + 1).astype(np.float32)
    act = np.tanh(obs[:, :, :cfg.act_dim] * 0.3 + 0.1 * rs.randn(n_traj, t_max, cfg.act_dim)).astype(np.float32)
    sc = S.Scaler(obs, act, True, dev)
    ds = DeviceWindowDataset(obs, act, lens, window=cfg.window, future_conditional=True, min_future_sep=5,
                             future_seq_len=cfg.goal_len, device=dev, scaler=sc)
    m = build_denoiser(cfg, dev, mode="precise", state_dict=synthetic_state_dict(cfg, 1))
    agent = BesoAgent(m, device=dev, sigma_min=0.005, sigma_max=1.0, window_size=cfg.window)
    agent.configure_training(lr=1e-3)
    agent.enable_data_parallel("nccl")
    m.train_math = os.environ.get("BESO_TRAIN_MATH", "fp32")

    def epoch(seed, max_steps):
        gen = torch.Generator().manual_seed(seed)                    # same permutation on every rank
        host_rng = np.random.RandomState(1000 * seed + rank)          # goal windows: per-rank draws
        losses = []
        for step, batch in enumerate(ds.batches(per_rank, shuffle=True, drop_last=True, generator=gen, rng=host_rng,
                                                rank=rank, world_size=world)):
            torch.manual_seed(10_000 * seed + 100 * step + rank)      # noise / sigma: per-rank draws
            losses.append(agent.train_step(batch))
            if step + 1 == max_steps:
                break
        return losses

    first = epoch(1, 6)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    timed = epoch(2, 10)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # replicas must be bit-identical: compare checksums and a strided sample of the raw and the EMA weights
    flat = torch.cat([p.detach().reshape(-1) for p in m.get_params()])
    ema = torch.cat([p.reshape(-1) for p in agent.ema_helper.shadow_params])
    probe = torch.cat([flat[::997], ema[::997], flat.double().sum().float().view(1), ema.double().sum().float().view(1)])
    gathered = [torch.empty_like(probe) for _ in range(world)]
    dist.all_gather(gathered, probe)
    identical = all(torch.equal(g, gathered[0]) for g in gathered)
    mean_first = torch.tensor([np.mean(first[:3])], device=dev); mean_last = torch.tensor([np.mean(timed[-3:])], device=dev)
    dist.all_reduce(mean_first); dist.all_reduce(mean_last)
    if rank == 0:
        print(json.dumps({"check": "dp agent train K256", "world": world, "global_batch": per_rank * world,
                          "train_math": m.train_math, "steps_timed": len(timed), "ms_per_step": t.item() / len(timed),
                          "samples_per_s": per_rank * world * len(timed) / (t.item() * 1e-3),
                          "replicas_identical": bool(identical), "loss_first3": mean_first.item() / world,
                          "loss_last3": mean_last.item() / world, "windows": len(ds)}))
        assert identical, "replicas diverged"
        assert mean_last.item() < mean_first.item(), "loss did not go down"
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
