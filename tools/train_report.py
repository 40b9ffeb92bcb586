"""GPU diagnostics of the training path: (a) per-parameter gradient error of each GEMM mode against the reference
goldens, (b) time and throughput of the cfg3 GEMM shapes per mode.

    python tools/train_report.py [errors|gemm|all]
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from beso_b200 import K256, _lib                               # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.synth import synthetic_state_dict              # noqa: E402
from beso_b200.training import flat_grad_views, loss_and_flat_grad  # noqa: E402

dev = torch.device("cuda:0")


def errors():
    from conftest import golden_weights, load_golden
    for name in ("loss_B256", "loss_K256"):
        cfg, meta, a = load_golden(name)
        m = build_denoiser(cfg, dev, mode="precise", state_dict=golden_weights(cfg, meta))
        m.train()
        g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in a.items()}
        for math in ("fp32", "bf16x2", "bf16"):
            m.train_math = math
            loss, flat = loss_and_flat_grad(m, g["state"], g["action"], g["goal"], g["noise"].clone(), g["sigma"])
            views = dict(zip([n for n, _ in m.named_parameters()], flat_grad_views(m, flat)))
            worst = []
            for n in [str(x) for x in a["grad_names"]]:
                fl = views[n].reshape(-1).cpu()
                got = fl if fl.numel() <= 4096 else fl[::97][:4096]
                want = a["grad::" + n]
                scale = float(want.abs().max())
                err = float((got - want).abs().max())
                bad = int(((got - want).abs() > 2e-3 * want.abs() + 1e-5 * scale + 2e-8).sum())
                worst.append((err / max(scale, 1e-30), n, err, scale, bad))
            worst.sort(reverse=True)
            print(f"{name} {math}: loss rel err {abs(float(loss) - float(a['loss'])) / float(a['loss']):.2e}; "
                  f"worst max|err|/max|grad| per tensor:")
            for w in worst[:6]:
                print(f"    {w[0]:.2e}  {w[1]}  (err {w[2]:.2e}, scale {w[3]:.2e}, outside tol: {w[4]})")


def gemm():
    lib = _lib.lib()
    m = build_denoiser(K256, dev, mode="precise", state_dict=synthetic_state_dict(K256, 1))
    m.refresh_weights()
    plan = m._plan
    M = 4096 * 22
    shapes = [("fwd qkv/proj  NT", M, 256, 256, 1, 1), ("fwd fc1       NT", M, 1024, 256, 1, 1), ("fwd fc2       NT", M, 256, 1024, 1, 1),
              ("dgrad fc2     NN", M, 1024, 256, 1, 0), ("dgrad fc1     NN", M, 256, 1024, 1, 0), ("dgrad d x d   NN", M, 256, 256, 1, 0),
              ("wgrad fc2     TN", 256, 1024, M, 0, 0), ("wgrad fc1     TN", 1024, 256, M, 0, 0), ("wgrad d x d   TN", 256, 256, M, 0, 0)]
    for name, Mm, N, K, ak, bk in shapes:
        A = torch.randn((Mm, K) if ak else (K, Mm), device=dev)
        Bm = torch.randn((N, K) if bk else (K, N), device=dev)
        out = torch.empty(Mm, N, device=dev)
        row = []
        for prec in (0, 1, 2):
            def run():
                _lib.check(lib.beso_debug_gemm(plan, A.data_ptr(), A.shape[1], ak, Bm.data_ptr(), Bm.shape[1], bk, out.data_ptr(), N,
                                               Mm, N, K, None, 0, prec, None), "gemm")
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            row.append(f"prec{prec} {ms:7.3f} ms {2.0 * Mm * N * K / ms / 1e9:7.1f} TF/s")
        gb = (Mm * K + N * K + Mm * N) * 4 / 1e9
        print(f"{name}  M={Mm} N={N} K={K}  {gb:5.2f} GB min traffic | " + " | ".join(row))
        torch.cuda.synchronize()
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        Am = A if ak else A.t()
        Bt = Bm.t() if bk else Bm
        for _ in range(3):
            torch.matmul(Am, Bt)
        a0.record()
        for _ in range(10):
            torch.matmul(Am, Bt)
        a1.record()
        torch.cuda.synchronize()
        print(f"        (library TF32 matmul of the same operands: {a0.elapsed_time(a1) / 10:.3f} ms)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("errors", "all"):
        errors()
    if what in ("gemm", "all"):
        gemm()
