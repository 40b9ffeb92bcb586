"""beso_b200 -- B200-native drop-in for the BESO score-based action-denoising hot path.

GCDenoiser (Karras pre-conditioning) -> DiffusionGPT (score-GPT forward) -> the
DDIM / Euler / Heun sample loop, as hand-written sm_100a CUDA kernels behind a
C-ABI shared library (include/beso_b200.h), mirrored on the host by the
reference's own Python interface.
"""
from .config import ModelConfig, K256, B256, T16, KITCHEN_CKPT, BLOCKPUSH_CKPT  # noqa: F401

__version__ = "0.1.0"
