"""One training-GEMM shape in a loop, for ncu.  python tools/gemm_one.py M N K a_kmajor b_kmajor prec [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import K256, _lib                               # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.synth import synthetic_state_dict              # noqa: E402

M, N, K, ak, bk, prec = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = torch.device("cuda:0")
lib = _lib.lib()
m = build_denoiser(K256, dev, mode="precise", state_dict=synthetic_state_dict(K256, 1))
m.refresh_weights()
A = torch.randn((M, K) if ak else (K, M), device=dev)
B = torch.randn((N, K) if bk else (K, N), device=dev)
out = torch.empty(M, N, device=dev)
for _ in range(reps):
    _lib.check(lib.beso_debug_gemm(m._plan, A.data_ptr(), A.shape[1], ak, B.data_ptr(), B.shape[1], bk, out.data_ptr(), N, M, N, K,
                                   None, 0, prec, None), "gemm")
torch.cuda.synchronize()
print("done")
