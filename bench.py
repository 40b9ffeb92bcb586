#!/usr/bin/env python
"""bench.py -- denoising-steps/s of the BESO sample loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode fast|precise] [--impl reference]

A "step" is one pass of the hot path over one batch: BASELINE config 2, the 50-step DDIM
``sample_loop`` over B = 512 sequences of the K256 score-GPT (obs 60, act 9, W 10, G 2, d 256,
L 4, H 4 -> 23 tokens) per GPU, i.e. 512 x 50 = 25,600 denoising steps per GPU per step, one
persistent kernel launch.  Weak scaling: every rank runs its own 512 sequences, no collective on
the data path.  value = denoising steps of all ranks / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, N_STEPS, SIGMA_MIN, SIGMA_MAX = 512, 50, 0.005, 1.0
METRIC, UNIT = "denoising-steps/sec (batch x steps)", "denoise-steps/s"
WORKLOAD = "cfg2: K256 score-GPT, 50-step DDIM sample_loop, batch 512 per GPU, synthetic obs/goal"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), "measured"
    return 1400.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def _nvml_loop(self):
        """Fast path: NVML through nvidia_ml_py (about 1 ms per sample); any failure falls back to nvidia-smi."""
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        bits = [("hw_slowdown", getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4))]
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop_evt.is_set():
            r = int(get_reasons(h))
            self.rows.append([str(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)), str(mx), str(N.nvmlDeviceGetPowerUsage(h) / 1000.0)] +
                             ["Active" if r & b else "Not Active" for _, b in bits])
            self._stop_evt.wait(0.01)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            pass
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def summary(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


def oracle_setup(batch):
    from beso_b200 import K256
    from beso_b200.synth import synthetic_inputs, synthetic_state_dict
    from oracle import beso_oracle as O
    cfg = K256
    oc = O.OracleCfg(obs_dim=cfg.obs_dim, act_dim=cfg.act_dim, window=cfg.window, goal_len=cfg.goal_len, d=cfg.d,
                     n_layers=cfg.n_layers, n_heads=cfg.n_heads, sigma_data=cfg.sigma_data)
    sd = O.as_module_params(synthetic_state_dict(cfg, seed=1))
    x = synthetic_inputs(cfg, batch, seed=2)
    sig = O.get_sigmas_exponential(N_STEPS, SIGMA_MIN, SIGMA_MAX)
    return O, oc, sd, x, sig


def time_oracle(batch, reps):
    """The reference's own CPU PyTorch path (oracle port, same ATen ops) on all host cores."""
    O, oc, sd, x, sig = oracle_setup(batch)
    times = []
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            O.sample_ddim(sd, oc, x["state"], x["noise"] * SIGMA_MAX, x["goal"], sig)
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args, out_fd):
    """--impl reference: the reference's CPU implementation of the same config, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_b = 128
    times = time_oracle(sample_b, args.warmup + args.steps)[args.warmup:]
    dt = sum(times) / len(times)
    value = sample_b * N_STEPS / dt
    sample = f"oracle port (torch {torch.__version__} CPU fp32) of {sample_b}/{BATCH} sequences x {N_STEPS} DDIM steps per step"
    emit(out_fd, {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def claim_stdout() -> int:
    """The contract is ONE JSON line on stdout.  Libraries also write there (NCCL prints its version line through
    C stdio, flushed at exit), so file descriptor 1 is pointed at stderr for the rest of the process and the JSON
    line goes to a private duplicate of the original stdout."""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def emit(fd: int, obj) -> None:
    os.write(fd, (json.dumps(obj) + "\n").encode())


def main():
    out_fd = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fast", "precise"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, out_fd)

    import ctypes as C
    from beso_b200 import K256, T16, _lib
    from beso_b200.denoiser import build_denoiser
    from beso_b200.sampling import ddim_coefficients, get_sigmas_exponential, sample_ddim
    from beso_b200.synth import synthetic_inputs, synthetic_state_dict

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    cfg = K256
    sd = synthetic_state_dict(cfg, seed=1)
    model = build_denoiser(cfg, dev, mode=args.mode, state_dict=sd)
    mode_id = model.resolved_mode()
    if mode_id == _lib.MODE_FAST:
        h = C.c_void_p()
        desc = _lib.ModelDesc.from_config(cfg)
        _lib.check(lib.beso_plan_create(C.byref(desc), local, C.byref(h)))
        if lib.beso_plan_rows_per_cta(h, _lib.MODE_FAST, cfg.window) <= 0:
            if args.mode == "fast":
                raise SystemExit("fast mode requested but not available in this build")
            model.mode, mode_id = "precise", _lib.MODE_PRECISE
        lib.beso_plan_destroy(h)
    mode_name = "fast" if mode_id == _lib.MODE_FAST else "precise"

    x = synthetic_inputs(cfg, BATCH, seed=2 + rank)
    sig = get_sigmas_exponential(N_STEPS, SIGMA_MIN, SIGMA_MAX)
    g_state, g_goal = x["state"].to(dev), x["goal"].to(dev)
    g_x = (x["noise"] * SIGMA_MAX).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        """Average device time per step (CUDA events, max over ranks) and this library's kernel launches inside the
        timed region (warm-up excluded)."""
        for _ in range(warmup):
            fn()
        barrier()
        n0 = lib.beso_kernel_launches()
        evs = []
        for _ in range(steps):
            flush.fill_(1)                                                 # evict L2 between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        timed.launches = int(lib.beso_kernel_launches() - n0)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    # ---- value: inputs resident in HBM, one persistent launch per step --------------------------
    clocks = ClockSampler(local)
    clocks.start()
    ms_step = timed(lambda: sample_ddim(model, g_state, g_x, g_goal, sig), args.steps, args.warmup)
    launches = timed.launches                                              # one persistent launch per step
    clock_summary = clocks.summary()
    steps_per_batch = BATCH * N_STEPS
    value = world * steps_per_batch / (ms_step * 1e-3)

    # ---- e2e: reference-facing C-ABI call with HOST buffers (H2D + kernel + D2H inside) ---------
    h_state, h_goal = x["state"].contiguous(), x["goal"].contiguous()
    h_x0 = (x["noise"] * SIGMA_MAX).contiguous()
    h_x = h_x0.clone()
    sig_arr = _lib.float_array(sig.tolist())
    coef_arr = _lib.float_array(ddim_coefficients(sig).reshape(-1).tolist())
    plan = model._plan
    stream = torch.cuda.current_stream(dev).cuda_stream

    def e2e_step():
        h_x.copy_(h_x0)
        _lib.check(lib.beso_sample_loop_host(plan, mode_id, _lib.SAMPLER_DDIM, sig_arr, N_STEPS + 1, coef_arr,
                                             h_state.data_ptr(), h_goal.data_ptr(), h_x.data_ptr(), BATCH, cfg.window,
                                             0, 0.0, C.c_void_p(stream)))

    for _ in range(max(3, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * steps_per_batch / e2e_s
    h2d = (h_state.numel() + h_goal.numel() + h_x.numel()) * 4
    d2h = h_x.numel() * 4

    # ---- roofline of the dominant (only) kernel ---------------------------------------------------
    sustained, burst, how = peaks()
    flops_launch = steps_per_batch * cfg.fwd_flops_per_seq()
    achieved = flops_launch / (ms_step * 1e-3) / 1e12
    traffic = None                 # DRAM bytes per launch from the committed ncu --set full capture of this command
    prof = os.path.join(ROOT, "profiles", "r1_fast_kernel.json")
    if mode_name == "fast" and os.path.exists(prof):
        try:
            d = json.load(open(prof))
            def _bytes(k):
                v, u = float(d[k]["value"].replace(",", "")), d[k]["unit"].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                "frac": achieved / sustained, "traffic": traffic,
                "kernel": "fast_sample_kernel<1>" if mode_name == "fast" else "simt_denoise_kernel",
                "peak_source": f"bf16_tflops_sustained of {how} (kernel runs for ms inside a long step)",
                "flops_per_launch": flops_launch,
                "note": ("fp16 tcgen05 (kind::f16) operands at the bf16 rate, fp32 accumulate in TMEM; peak = measured bf16 dense" if mode_name == "fast" else
                         "precise mode computes on the fp32 FMA pipe; the tensor peak is quoted for comparability")}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16" if mode_name == "fast" else "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "mode": mode_name, "sampler": "ddim", "n_sampling_steps": N_STEPS,
                      "batch_per_gpu": BATCH, "tokens_per_seq": cfg.n_tokens(), "l2": "flushed between timed steps (256 MiB write)",
                      "weights": "synthetic N(0,0.02) seed 1", "parallelism": f"replicas x{world}, no collective"},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": launches, "clocks": clock_summary, "roofline": roofline}

    if rank == 0 and not args.no_extras:
        extras = {}
        # north-star shape: one fused forward, B = 4096, 16 tokens (T16) and 23 tokens (K256)
        for name, c in (("fwd_T16_b4096", T16), ("fwd_K256_b4096", K256)):
            mm = build_denoiser(c, dev, mode=mode_name, state_dict=synthetic_state_dict(c, seed=3))
            xi = {k: v.to(dev) for k, v in synthetic_inputs(c, 4096, seed=4).items()}
            if dist is None:
                ms = timed(lambda: mm(xi["state"], xi["action"], xi["goal"], xi["sigma"]), 10, 3)
            else:
                continue
            tf = 4096 * c.fwd_flops_per_seq() / (ms * 1e-3) / 1e12
            extras[name] = {"ms": ms, "denoise_steps_per_s": 4096 / (ms * 1e-3), "tflops": tf,
                            "frac_of_burst_peak": tf / burst, "frac_of_sustained_peak": tf / sustained}
        # parity of the benched mode on a slice of the benched workload (oracle = checker only)
        from oracle import beso_oracle as O
        Oo, oc, osd, ox, osig = oracle_setup(8)
        with torch.no_grad():
            want = Oo.sample_ddim(osd, oc, ox["state"], ox["noise"] * SIGMA_MAX, ox["goal"], osig)
        got = sample_ddim(model, ox["state"].to(dev), (ox["noise"] * SIGMA_MAX).to(dev), ox["goal"].to(dev), sig).cpu()
        err = (got - want).abs()
        extras["parity_vs_oracle_ddim50_b8"] = {
            "max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()),
            "frac_within_rtol1e-3_atol1e-5": float((err <= 1e-5 + 1e-3 * want.abs()).float().mean())}
        out["extras"] = extras

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sample_b = 128
        times = time_oracle(sample_b, 3)[1:]
        dt = min(times)
        out["cpu_baseline"] = {"value": sample_b * N_STEPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{sample_b}/{BATCH} sequences x {N_STEPS} DDIM steps, best of 2 after 1 warm-up, "
                                         f"oracle port on torch {torch.__version__} CPU fp32"}
    if rank == 0:
        emit(out_fd, out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
