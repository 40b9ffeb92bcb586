"""ctypes binding of libbeso_b200.so (include/beso_b200.h).

The library is built in-tree by ``beso_b200/csrc/Makefile`` (nvcc, sm_100a only) into
``beso_b200/lib/``.  There is no CPU fallback: if the library cannot be loaded every product
entry point raises ``BesoLibraryError``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbeso_b200.so")
CSRC = os.path.join(_HERE, "csrc")

MODE_PRECISE, MODE_FAST, MODE_SIMT = 0, 1, 2
SAMPLER_DDIM, SAMPLER_EULER, SAMPLER_HEUN, SAMPLER_EULER_ANCESTRAL, SAMPLER_DPMPP_2M, SAMPLER_TWO_STAGE, SAMPLER_LMS = 0, 1, 2, 3, 4, 5, 6
FLAG_UNCOND, FLAG_CFG, FLAG_INNER, FLAG_PRED_LAST, FLAG_TRAIN_FAST = 1, 2, 4, 8, 16
FLAG_TRAIN_TF32 = FLAG_TRAIN_FAST
FLAG_TRAIN_SPLIT2 = 32
SAMPLER_IDS = {"ddim": SAMPLER_DDIM, "euler": SAMPLER_EULER, "heun": SAMPLER_HEUN, "euler_ancestral": SAMPLER_EULER_ANCESTRAL,
               "dpmpp_2m": SAMPLER_DPMPP_2M, "two_stage": SAMPLER_TWO_STAGE,
               "lms": SAMPLER_LMS}
MODE_IDS = {"precise": MODE_PRECISE, "fast": MODE_FAST, "simt": MODE_SIMT}

EXPORTS = [
    "beso_last_error", "beso_abi_version", "beso_param_count", "beso_param_numel", "beso_param_total",
    "beso_plan_create", "beso_plan_destroy", "beso_plan_pack_weights", "beso_plan_select_weights", "beso_plan_set_params",
    "beso_denoise_fwd", "beso_sample_loop", "beso_sample_loop_noise", "beso_sample_loop_scaled", "beso_denoise_fwd_host", "beso_sample_loop_host",
    "beso_loss_fwd_bwd", "beso_loss_fwd_bwd_dropout", "beso_loss_fwd_bwd_dp", "beso_debug_gemm", "beso_comm_unique_id", "beso_comm_init", "beso_comm_destroy",
    "beso_allreduce_grads", "beso_kernel_launches", "beso_plan_rows_per_cta", "beso_device_sm_count",
    "beso_debug_set_trace", "beso_debug_set_timeline", "beso_debug_mma_rate", "beso_debug_set_precise_layout",
    "beso_opt_create", "beso_opt_destroy", "beso_opt_total", "beso_opt_step", "beso_window_gather",
]


class BesoLibraryError(RuntimeError):
    pass


class ModelDesc(C.Structure):
    """struct beso_model_desc"""
    _fields_ = [("obs_dim", C.c_int32), ("act_dim", C.c_int32), ("window", C.c_int32),
                ("goal_len", C.c_int32), ("d", C.c_int32), ("n_layers", C.c_int32),
                ("n_heads", C.c_int32), ("linear_output", C.c_int32), ("goal_conditioned", C.c_int32),
                ("sigma_data", C.c_float)]

    @classmethod
    def from_config(cls, cfg) -> "ModelDesc":
        return cls(cfg.obs_dim, cfg.act_dim, cfg.window, cfg.goal_len, cfg.d, cfg.n_layers, cfg.n_heads,
                   int(cfg.linear_output), int(cfg.goal_conditioned), float(cfg.sigma_data))


class IoScaling(C.Structure):
    """beso_io_scaling of include/beso_b200.h."""
    _fields_ = [("in_table", C.c_void_p), ("goal_keep", C.c_void_p), ("out_clip", C.c_void_p), ("out_table", C.c_void_p),
                ("unscaled_out", C.c_void_p)]


class DropoutMasks(C.Structure):
    """beso_dropout_masks of include/beso_b200.h."""
    _fields_ = [("embed", C.c_void_p), ("attn", C.POINTER(C.c_void_p)), ("resid_attn", C.POINTER(C.c_void_p)),
                ("resid_mlp", C.POINTER(C.c_void_p))]


_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile libbeso_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise BesoLibraryError("building libbeso_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def _declare(lib):
    fp, vp, i32, u32, f32 = C.POINTER(C.c_float), C.c_void_p, C.c_int, C.c_uint32, C.c_float
    pd = C.POINTER(ModelDesc)
    lib.beso_last_error.restype = C.c_char_p
    lib.beso_last_error.argtypes = []
    lib.beso_abi_version.restype = i32
    lib.beso_param_count.argtypes = [pd]
    lib.beso_param_numel.argtypes = [pd, i32]
    lib.beso_param_numel.restype = C.c_int64
    lib.beso_param_total.argtypes = [pd]
    lib.beso_param_total.restype = C.c_int64
    lib.beso_plan_create.argtypes = [pd, i32, C.POINTER(vp)]
    lib.beso_plan_destroy.argtypes = [vp]
    lib.beso_plan_pack_weights.argtypes = [vp, i32, C.POINTER(vp), i32, vp]
    lib.beso_plan_select_weights.argtypes = [vp, i32]
    lib.beso_plan_set_params.argtypes = [vp, i32, C.POINTER(vp), i32]
    lib.beso_denoise_fwd.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32, u32, f32, vp]
    lib.beso_sample_loop.argtypes = [vp, i32, i32, fp, i32, fp, vp, vp, vp, i32, i32, u32, f32, vp]
    lib.beso_sample_loop_noise.argtypes = [vp, i32, i32, fp, i32, fp, vp, vp, vp, vp, i32, i32, u32, f32, vp]
    lib.beso_sample_loop_scaled.argtypes = [vp, i32, i32, fp, i32, fp, vp, vp, vp, vp, C.POINTER(IoScaling), i32, i32, u32, f32, vp]
    lib.beso_denoise_fwd_host.argtypes = [vp, i32, vp, vp, vp, vp, vp, i32, i32, u32, f32, vp]
    lib.beso_sample_loop_host.argtypes = [vp, i32, i32, fp, i32, fp, vp, vp, vp, i32, i32, u32, f32, vp]
    lib.beso_loss_fwd_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, u32, vp]
    lib.beso_loss_fwd_bwd_dropout.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(DropoutMasks), vp, vp, i32, u32, vp]
    lib.beso_loss_fwd_bwd_dp.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(DropoutMasks), vp, f32, vp, vp, i32, u32, vp]
    lib.beso_debug_gemm.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp]
    lib.beso_comm_unique_id.argtypes = [C.c_char_p]
    lib.beso_comm_init.argtypes = [i32, i32, C.c_char_p, i32, C.POINTER(vp)]
    lib.beso_comm_destroy.argtypes = [vp]
    lib.beso_allreduce_grads.argtypes = [vp, vp, C.c_size_t, f32, vp]
    lib.beso_kernel_launches.restype = C.c_int64
    lib.beso_plan_rows_per_cta.argtypes = [vp, i32, i32]
    lib.beso_device_sm_count.argtypes = [i32]
    lib.beso_debug_set_precise_layout.argtypes = [i32]
    lib.beso_debug_set_trace.argtypes = [vp]
    lib.beso_debug_set_timeline.argtypes = [vp]
    lib.beso_debug_mma_rate.argtypes = [vp, vp, i32, vp]
    lib.beso_opt_create.argtypes = [i32, i32, C.POINTER(vp), C.POINTER(C.c_longlong), C.POINTER(vp)]
    lib.beso_opt_destroy.argtypes = [vp]
    lib.beso_opt_total.argtypes = [vp]
    lib.beso_opt_total.restype = C.c_longlong
    lib.beso_opt_step.argtypes = [vp, vp, vp, vp, vp, f32, f32, f32, f32, f32, i32, f32, f32, vp]
    lib.beso_window_gather.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, i32, vp]


def lib():
    """Returns the loaded library; raises BesoLibraryError (never falls back) if unavailable."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise BesoLibraryError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "or `make -C beso_b200/csrc`.  beso_b200 has no CPU or PyTorch fallback.")
            try:
                import torch  # noqa: F401  (loads libcudart / libnccl that the library links against)
                handle = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
            except OSError as e:  # pragma: no cover
                raise BesoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
            _declare(handle)
            if handle.beso_abi_version() != 1:
                raise BesoLibraryError("libbeso_b200.so ABI version mismatch")
            _lib = handle
        return _lib


def check(rc: int, what: str = "") -> int:
    if rc < 0:
        msg = lib().beso_last_error().decode()
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        raise BesoLibraryError(f"{what}: {msg} (code {rc})")
    return rc


def float_array(values):
    arr = (C.c_float * len(values))(*[float(v) for v in values])
    return arr
