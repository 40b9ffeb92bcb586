"""GPU diagnostics: tcgen05.mma completion cycles for a few instruction mixes (64 MMAs each), alone and
with concurrent bulk-copy (TMA) writes and / or LDS/STS traffic in the same SM."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from beso_b200 import _lib  # noqa: E402

names = ["N256 no-commit", "N256 commit/4", "N128 no-commit", "N128 commit/4", "N256 alt-D", "N256 commit/1",
         "N64 no-commit", "N192 no-commit", "N256 commit/16", "N128 commit/16"]
src = torch.zeros(6 << 20, dtype=torch.uint8, device="cuda")
for mode, label in ((0, "quiet SM"), (1, "+ bulk-copy ring"), (2, "+ LDS/STS warps"), (3, "+ both")):
    out = torch.zeros(24, dtype=torch.int64, device="cuda")
    for rep in range(2):
        _lib.check(_lib.lib().beso_debug_mma_rate(C.c_void_p(out.data_ptr()), C.c_void_p(src.data_ptr()), mode, None))
        torch.cuda.synchronize()
    v = out.cpu().tolist()
    total = sum(v[2 * i + 1] for i in range(10))
    print(f"== {label}: per-MMA cycles " + "  ".join(f"{names[i]}={v[2 * i + 1] / 64:.0f}" for i in range(10)) +
          (f"  | copies {v[20]} = {v[20] * 16384 / max(total, 1):.1f} B/cycle" if mode & 1 else ""))
