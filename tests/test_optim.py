"""Fused AdamW + EMA step against torch.optim.AdamW and the reference's EMA formula.

Tolerance: fp32, same operation order as torch's single-tensor AdamW; rtol 2e-6 per step (torch's CUDA
kernels may contract multiply-adds differently), checked over several steps with a StepLR schedule.
"""
import pytest
import torch

from beso_b200 import _lib
from beso_b200.optim import ExponentialMovingAverage, FusedAdamW, ema_decay_at


def test_ema_decay_schedule_matches_reference_formula():
    # beso/networks/ema_helper/ema.py:46-50: num_updates is incremented first
    assert ema_decay_at(0.999, None) == 0.999
    assert ema_decay_at(0.999, 1) == pytest.approx(2 / 11)
    assert ema_decay_at(0.999, 10) == pytest.approx(11 / 20)
    assert ema_decay_at(0.999, 100000) == 0.999


def test_library_exports_optimizer_entry_points():
    lib = _lib.lib()
    for name in ("beso_opt_create", "beso_opt_destroy", "beso_opt_total", "beso_opt_step"):
        assert hasattr(lib, name)


def test_fused_adamw_refuses_cpu_parameters():
    p = torch.nn.Parameter(torch.zeros(4))
    with pytest.raises(_lib.BesoLibraryError):
        FusedAdamW([p], lr=1e-3)


def _reference_ema_update(shadow, params, decay, num_updates):
    num_updates += 1
    d = min(decay, (1 + num_updates) / (10 + num_updates))
    for s, p in zip(shadow, params):
        s.sub_((1.0 - d) * (s - p))
    return num_updates


@pytest.mark.gpu
@pytest.mark.parametrize("weight_decay", [0.01, 0.0])
def test_fused_step_matches_torch_adamw_and_reference_ema(weight_decay):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(1, 13, 256), (256, 60), (256,), (1024, 256), (1024,), (256, 1024), (9, 256), (9,), (5,), (3, 7)]
    ours = [torch.nn.Parameter((torch.randn(s, generator=g) * 0.02).to(dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt_ref = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), weight_decay=weight_decay, foreach=False, fused=False)
    sch_ref = torch.optim.lr_scheduler.StepLR(opt_ref, step_size=2, gamma=0.5)
    opt = FusedAdamW(ours, lr=1e-3, betas=(0.9, 0.999), weight_decay=weight_decay)
    sch = torch.optim.lr_scheduler.StepLR(opt, step_size=2, gamma=0.5)
    ema = ExponentialMovingAverage(ours, decay=0.999)
    opt.attach_ema(ema)
    shadow_ref = [p.detach().clone() for p in ref]
    n_upd = 0
    launches0 = _lib.lib().beso_kernel_launches()
    for step in range(7):
        grads = [(torch.randn(s, generator=g) * (0.1 if step % 2 else 1e-3)).to(dev) for s in shapes]
        for p, q, gr in zip(ours, ref, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        v0 = ours[0]._version
        opt_ref.step(); sch_ref.step()
        with torch.no_grad():
            n_upd = _reference_ema_update(shadow_ref, ref, 0.999, n_upd)
        opt.step(); sch.step()
        ema.update(ours)
        assert ours[0]._version > v0                      # the denoiser re-packs on version change
        for p, q in zip(ours, ref):
            torch.testing.assert_close(p, q, rtol=2e-6 * (step + 1), atol=1e-9)
        for s, r in zip(ema.shadow_params, shadow_ref):
            torch.testing.assert_close(s, r, rtol=2e-6 * (step + 1), atol=1e-9)
    assert _lib.lib().beso_kernel_launches() - launches0 == 7     # one launch per step
    assert ema.num_updates == n_upd
    with pytest.raises(_lib.BesoLibraryError):
        ema.update(ours)                                    # nothing pending: no stand-alone path


@pytest.mark.gpu
def test_fused_step_with_flat_gradient_and_scale():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    shapes = [(64, 32), (64,), (7,)]
    ours = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt_ref = torch.optim.AdamW(ref, lr=3e-4, foreach=False, fused=False)
    opt = FusedAdamW(ours, lr=3e-4)
    flat = torch.randn(sum(p.numel() for p in ours), generator=g).to(dev)
    off = 0
    for q in ref:
        q.grad = (flat[off:off + q.numel()] * 0.5).view_as(q).clone()
        off += q.numel()
    opt_ref.step()
    opt.step(flat_grad=flat, grad_scale=0.5)                # e.g. 1 / world after an all-reduce(sum)
    for p, q in zip(ours, ref):
        torch.testing.assert_close(p, q, rtol=2e-6, atol=1e-9)
