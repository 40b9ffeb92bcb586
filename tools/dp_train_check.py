"""Multi-GPU check of BASELINE config 4 (data-parallel training step): run under torchrun, one rank per GPU.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py

Every rank computes loss + flat gradient of its shard of a global K256 batch (1024 sequences per rank),
the flat buffer is all-reduced with ncclAllReduce through the C ABI and scaled by 1/N; rank 0 also computes
the full global batch alone and the two gradients must agree.  Prints timings (device, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beso_b200 import K256                                          # noqa: E402
from beso_b200.denoiser import build_denoiser                       # noqa: E402
from beso_b200.dist import FlatGradAllReduce, shard_batch           # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402
from beso_b200.training import loss_and_flat_grad                   # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg, per = K256, 1024
    m = build_denoiser(cfg, dev, mode="precise", state_dict=synthetic_state_dict(cfg, 1))
    m.train()
    x = {k: v.to(dev) for k, v in synthetic_inputs(cfg, per * world, seed=7, sigma_min=0.05).items()}
    sl = shard_batch(per * world, rank, world)
    ar = FlatGradAllReduce("nccl", device=local)

    def step():
        loss, flat = loss_and_flat_grad(m, x["state"][sl], x["clean"][sl], x["goal"][sl], x["noise"][sl], x["sigma"][sl])
        ar(flat)
        return loss, flat

    for _ in range(3):
        step()
    dist.barrier(); torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(5):
        loss, flat = loss_and_flat_grad(m, x["state"][sl], x["clean"][sl], x["goal"][sl], x["noise"][sl], x["sigma"][sl])
    e[1].record()
    for _ in range(5):
        ar(flat)
    e[2].record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    loss, flat = step()
    ok, err = True, 0.0
    if rank == 0:
        _, full = loss_and_flat_grad(m, x["state"], x["clean"], x["goal"], x["noise"], x["sigma"])
        err = float((flat - full).abs().max() / full.abs().max())
        ok = err < 2e-3
        print(json.dumps({"check": "dp_train K256", "world": world, "global_batch": per * world,
                          "compute_ms_per_step": float(t[0]), "allreduce_ms": float(t[1]),
                          "grad_floats": flat.numel(), "rel_err_vs_single_gpu_full_batch": err, "ok": ok,
                          "samples_per_s": per * world / ((float(t[0]) + float(t[1])) * 1e-3)}))
    ar.close()
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
