#!/usr/bin/env python
"""bench.py -- denoising-steps/s of the BESO sample loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode precise|fast|simt] [--impl reference]

The default (and the library's default) mode is "precise": fp32-equivalent arithmetic that meets the north-star tolerance
(rtol 1e-3 / atol 1e-5 against the fp32 reference), on the tcgen05 tensor cores with split fp16 operands.  "fast" (single-pass
fp16 operands, outside that tolerance) is timed in the same run and reported under extras.fast.

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


CPU_PROTOCOL = ("oracle port (torch {tv} CPU fp32, the reference's ATen ops in the reference's order) of the FULL workload "
                "(512 sequences x 50 DDIM steps) on {cores} host threads, mean of {k} timed passes after {w} warm-up")


def run_reference(args, out_fd):
    """--impl reference: the reference's CPU implementation of the same config (full batch), rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    warm = max(1, args.warmup)
    times = time_oracle(BATCH, warm + args.steps)[warm:]
    dt = sum(times) / len(times)
    value = BATCH * N_STEPS / dt
    sample = CPU_PROTOCOL.format(tv=torch.__version__, cores=cores, k=len(times), w=warm)
    emit(out_fd, {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
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


def committed_traffic(mode_name):
    """DRAM bytes per launch of the benched kernel from this round's committed `ncu --set full` capture of this command
    (profiles/r2_<mode>_kernel.json: dram__bytes_read.sum + dram__bytes_write.sum).  ncu cannot run inside a timed
    bench; the file is re-captured whenever the kernel changes and names the commit it was taken at."""
    prof = os.path.join(ROOT, "profiles", f"r2_{mode_name}_kernel.json")
    if not os.path.exists(prof):
        return None, None
    try:
        d = json.load(open(prof))

        def _bytes(k):
            v, u = float(str(d[k]["value"]).replace(",", "")), d[k]["unit"].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        committed_traffic.tensor_active = float(str(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]["value"]).replace(",", ""))
        return _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum"), os.path.relpath(prof, ROOT)
    except Exception:
        return None, None


committed_traffic.tensor_active = None


def main():
    out_fd = claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fast", "precise", "simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, out_fd)

    import ctypes as C
    from beso_b200 import B256, K256, T16, _lib
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
    model.refresh_weights()
    mode_id = model.resolved_mode()
    if mode_id == _lib.MODE_FAST and not model.fast_supported():
        raise SystemExit("fast mode requested but this shape is not supported by the tensor-core kernel")
    mode_name = {_lib.MODE_FAST: "fast", _lib.MODE_PRECISE: "precise", _lib.MODE_SIMT: "simt"}[mode_id]

    x = synthetic_inputs(cfg, BATCH, seed=2 + rank)
    sig = get_sigmas_exponential(N_STEPS, SIGMA_MIN, SIGMA_MAX)
    g_state, g_goal = x["state"].to(dev), x["goal"].to(dev)
    g_x = (x["noise"] * SIGMA_MAX).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, sync_ranks=True):
        """Average device time per step (CUDA events on the launching stream, max over ranks) and this library's kernel
        launches inside the timed region (warm-up excluded)."""
        for _ in range(warmup):
            fn()
        barrier() if sync_ranks else torch.cuda.synchronize(dev)
        n0 = lib.beso_kernel_launches()
        evs = []
        for _ in range(steps):
            flush.fill_(1)                                                 # evict L2 between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier() if sync_ranks else torch.cuda.synchronize(dev)
        timed.launches = int(lib.beso_kernel_launches() - n0)
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if dist is not None and sync_ranks:
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
    # The timed region is K isolated launches of a few ms with an L2 flush between them, at the maximum SM clock (the
    # clocks record shows it): the denominator is the BURST measured peak; the sustained figure is given beside it.
    sustained, burst, how = peaks()
    flops_launch = steps_per_batch * cfg.fwd_flops_per_seq()
    achieved = flops_launch / (ms_step * 1e-3) / 1e12
    traffic, traffic_src = committed_traffic(mode_name)
    kernel_name = {"fast": "fast_sample_kernel<1,false,1,false> (fp16 tcgen05, single pass)",
                   "precise": "fast_sample_kernel<1,false,1,true,64,2> (split fp16 operands on tcgen05, 128-row tiles: A_hi W_hi + A_lo W_hi + A_hi W_lo "
                              "per product, the lo image of the LayerNorm output read as a tensor-memory A operand)",
                   "simt": "simt_denoise_kernel (fp32 FMA)"}[mode_name]
    roofline = {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                "frac": achieved / burst, "frac_of_sustained": achieved / sustained, "peak_sustained": sustained,
                "traffic": traffic, "traffic_source": traffic_src,
                # tensor-pipe active cycles (% of elapsed) of the same committed ncu capture: what the pipe was busy with,
                # padding rows and the precise mode's extra MMAs included
                "tensor_pipe_active_pct_of_elapsed_ncu": committed_traffic.tensor_active,
                "kernel": kernel_name,
                "peak_source": f"bf16_tflops (burst) of {how} MEASURED_PEAKS.json: the kernel is timed alone, ms-long launches at max clock",
                "flops_per_launch": flops_launch,
                # tile rows used by this workload: 512 sequences of 23 tokens go out as 128 tiles of 4 sequences (92 of 128 rows)
                # in both modes (one wave over the 148 SMs)
                "tensor_flops_executed_per_algorithmic": {"fast": 1.0 * 128 / 92, "precise": 3.0 * 128 / 92, "simt": 0.0}[mode_name],
                "note": "algorithmic FLOPs (SURVEY.md 8d: every product counted once, no padding) / device time.  The precise mode "
                        "runs every product as fp16 hi / lo operand images, three M=128 MMAs per product (the lo.lo term is below "
                        "2^-22) where the fp16 mode runs one: 3 tensor FLOPs per algorithmic FLOP (x 128/92 row padding at 4 "
                        "sequences of 23 tokens per tile), so the algorithmic fraction of the precise mode is bounded by 1/3 of "
                        "the fp16 mode's; its tensor-pipe activity is in profiles/r2_precise_kernel.md"}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f16" if mode_name == "fast" else "f32", "data": "synthetic",
           "tolerance": {"fast": "rtol 3e-3 / atol 3e-3 (outside the north-star tolerance; opt-in mode)",
                         "precise": "rtol 1e-3 / atol 1e-5 vs the fp32 reference (north-star tolerance): met, see extras.precise.parity",
                         "simt": "rtol 1e-3 / atol 1e-5 vs the fp32 reference (north-star tolerance)"}[mode_name],
           "config": {"workload": WORKLOAD, "mode": mode_name, "sampler": "ddim", "n_sampling_steps": N_STEPS,
                      "batch_per_gpu": BATCH, "tokens_per_seq": cfg.n_tokens(), "l2": "flushed between timed steps (256 MiB write)",
                      "weights": "synthetic N(0,0.02) seed 1", "parallelism": f"replicas x{world}, no collective on the sampling path"},
           "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
           "gpu_launches": launches, "clocks": clock_summary, "roofline": roofline}

    extras = {}
    if rank == 0 and not args.no_extras:
        from oracle import beso_oracle as O
        Oo, oc, osd, ox, osig = oracle_setup(8)
        with torch.no_grad():
            want = Oo.sample_ddim(osd, oc, ox["state"], ox["noise"] * SIGMA_MAX, ox["goal"], osig)

        def parity(m):
            got = sample_ddim(m, ox["state"].to(dev), (ox["noise"] * SIGMA_MAX).to(dev), ox["goal"].to(dev), sig).cpu()
            err = (got - want).abs()
            return {"max_abs_err": float(err.max()), "mean_abs_err": float(err.mean()),
                    "frac_within_rtol1e-3_atol1e-5": float((err <= 1e-5 + 1e-3 * want.abs()).float().mean())}

        # both arithmetic modes of the tensor-core kernel on the bench workload and on the north-star forward shapes
        # (one fused forward, B = 4096, 16 tokens (T16) and 23 tokens (K256)); rank 0 only, not synchronised with the
        # other ranks (they idle at the barrier below)
        models = {m_: (model if m_ == mode_name else build_denoiser(cfg, dev, mode=m_, state_dict=sd)) for m_ in ("fast", "precise")}
        for mname in ("fast", "precise"):
            mm = models[mname]
            e = {"parity_vs_oracle_ddim50_b8": parity(mm)}
            if mname == "precise":
                # 8 sequences would take the stacked 64-row tiles; the timed workload (512 sequences) runs the 128-row
                # layout: check that one too (test hook of the C ABI), then hand the choice back to the launch
                lib.beso_debug_set_precise_layout(2)
                e["parity_vs_oracle_ddim50_b8_128row_tiles"] = parity(mm)
                lib.beso_debug_set_precise_layout(0)
            if mname != mode_name:
                ms = timed(lambda: sample_ddim(mm, g_state, g_x, g_goal, sig), max(3, args.steps // 4), 3, sync_ranks=False)
                e.update({"ms_per_step": ms, "value": steps_per_batch / (ms * 1e-3), "unit": UNIT,
                          "tflops": flops_launch / (ms * 1e-3) / 1e12, "frac_of_burst_peak": flops_launch / (ms * 1e-3) / 1e12 / burst})
            else:
                e.update({"ms_per_step": ms_step, "value": steps_per_batch / (ms_step * 1e-3), "unit": UNIT,
                          "tflops": achieved, "frac_of_burst_peak": achieved / burst})
            for name, c in (("fwd_T16_b4096", T16), ("fwd_K256_b4096", K256)):
                fm = build_denoiser(c, dev, mode=mname, state_dict=synthetic_state_dict(c, seed=3))
                xi = {k: v.to(dev) for k, v in synthetic_inputs(c, 4096, seed=4).items()}
                ms = timed(lambda: fm(xi["state"], xi["action"], xi["goal"], xi["sigma"]), 10, 3, sync_ranks=False)
                tf = 4096 * c.fwd_flops_per_seq() / (ms * 1e-3) / 1e12
                e[name] = {"ms": ms, "denoise_steps_per_s": 4096 / (ms * 1e-3), "tflops": tf,
                           "frac_of_burst_peak": tf / burst, "frac_of_sustained_peak": tf / sustained}
            extras[mname] = e
        # rollout latency: predict()-shaped call, batch 1, 10-step DDIM, one launch
        xr = {k: v.to(dev) for k, v in synthetic_inputs(cfg, 1, seed=5).items()}
        sig10 = get_sigmas_exponential(10, SIGMA_MIN, SIGMA_MAX)
        extras["rollout_b1_ddim10_ms"] = {
            mname: timed(lambda: sample_ddim(models[mname], xr["state"], xr["noise"], xr["goal"], sig10), 20, 5, sync_ranks=False)
            for mname in ("fast", "precise")}
        # the reference's REAL model shapes (frozen configs beside the shipped checkpoints): kitchen d = 360 / 6 heads of
        # 60 / 6 layers (the 384-column geometry of the tensor-core kernel) and block-push d = 240 / 12 heads of 20 / 4
        # layers (padded 256 / 32); one forward at batch 4096 and the predict()-shaped rollout call at batch 1
        from beso_b200 import BLOCKPUSH_CKPT, KITCHEN_CKPT
        ck = {}
        for label, c in (("kitchen_d360", KITCHEN_CKPT), ("blockpush_d240", BLOCKPUSH_CKPT)):
            csd = synthetic_state_dict(c, seed=7)
            xi = {k: v.to(dev) for k, v in synthetic_inputs(c, 4096, seed=8).items()}
            x1 = {k: v.to(dev) for k, v in synthetic_inputs(c, 1, seed=9).items()}
            row = {}
            for mname in ("fast", "precise"):
                fm = build_denoiser(c, dev, mode=mname, state_dict=csd)
                assert fm.fast_supported()                                # both shapes run the tensor-core kernel
                ms = timed(lambda: fm(xi["state"], xi["action"], xi["goal"], xi["sigma"]), 10, 3, sync_ranks=False)
                tf = 4096 * c.fwd_flops_per_seq() / (ms * 1e-3) / 1e12
                ms1 = timed(lambda: sample_ddim(fm, x1["state"], x1["noise"], x1["goal"], sig10), 10, 3, sync_ranks=False)
                row[mname] = {"fwd_b4096_ms": ms, "denoise_steps_per_s": 4096 / (ms * 1e-3), "tflops": tf,
                              "frac_of_burst_peak": tf / burst, "rollout_b1_ddim10_ms": ms1}
            ck[label] = row
        extras["checkpoint_shapes"] = ck
        # training step (BASELINE config 3: block-push shape, batch 4096): fused loss + backward, tcgen05 GEMMs
        from beso_b200.training import loss_and_flat_grad
        tm = build_denoiser(B256, dev, mode="precise", state_dict=synthetic_state_dict(B256, 41))
        tm.train()
        tg = {k: v.to(dev) for k, v in synthetic_inputs(B256, 4096, seed=42, sigma_min=0.05).items()}
        tr = {}
        for math in ("fp32", "bf16x2", "bf16"):
            tm.train_math = math
            ms = timed(lambda: loss_and_flat_grad(tm, tg["state"], tg["clean"], tg["goal"], tg["noise"], tg["sigma"]), 5, 2, sync_ranks=False)
            tr[math] = {"ms_fwd_bwd": ms, "samples_per_s": 4096 / (ms * 1e-3),
                        "tflops_algorithmic": 3 * 4096 * B256.fwd_flops_per_seq() / (ms * 1e-3) / 1e12}
        extras["train_cfg3_b4096"] = tr
        del tm, tg

    # ---- data-parallel training leg (BASELINE config 4): the one collective of the design ------------------------
    if world > 1 and not args.no_extras:
        from beso_b200.dist import FlatGradAllReduce
        from beso_b200.optim import FusedAdamW
        from beso_b200.training import loss_and_flat_grad
        per_rank = 1024
        dm = build_denoiser(K256, dev, mode="precise", state_dict=synthetic_state_dict(K256, 1))
        dm.train()
        dx = {k: v.to(dev) for k, v in synthetic_inputs(K256, per_rank, seed=1000 + rank).items()}
        sync = FlatGradAllReduce("nccl", device=local)                     # the repo's own NCCL communicator
        opt = FusedAdamW(list(dm.get_params()), lr=1e-4)
        a = (dx["state"], dx["clean"], dx["goal"], dx["noise"], dx["sigma"])

        def compute_only():
            return loss_and_flat_grad(dm, *a)

        def overlapped():
            loss, flat = loss_and_flat_grad(dm, *a, grad_sync=sync)        # per-block all-reduce behind the backward
            return flat

        def sequential():
            loss, flat = loss_and_flat_grad(dm, *a)
            return sync(flat)                                              # one all-reduce after the backward

        def full_step():
            flat = overlapped()
            opt.step(flat_grad=flat)

        # exchange check against an independent collective: torch.distributed's all-reduce of the local gradients
        _, local_flat = compute_only()
        ref = local_flat.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref.mul_(1.0 / world)
        got = overlapped()
        scale = float(ref.abs().max())
        err_overlap = float((got - ref).abs().max()) / scale
        ms_c = timed(compute_only, 5, 2)
        ms_o = timed(overlapped, 5, 2)
        ms_s = timed(sequential, 5, 2)
        ms_f = timed(full_step, 5, 2)
        if rank == 0:
            n_grad = int(local_flat.numel())
            extras["dp_train_cfg4"] = {
                "per_rank_batch": per_rank, "global_batch": per_rank * world, "grad_floats": n_grad,
                "ms_compute_only": ms_c, "ms_with_overlapped_allreduce": ms_o, "ms_with_allreduce_after_backward": ms_s,
                "ms_exposed_comm_overlapped": ms_o - ms_c, "ms_exposed_comm_sequential": ms_s - ms_c,
                "ms_full_step_incl_adamw_ema": ms_f, "samples_per_s": per_rank * world / (ms_f * 1e-3),
                "allreduce_vs_torch_distributed_max_err_over_scale": err_overlap,
                "collective": "ncclAllReduce per transformer block (beso_loss_fwd_bwd_dp) on the repo's own communicator, "
                              f"{n_grad * 4 / 1e6:.1f} MB fp32 per step"}
        sync.close()

    if rank == 0 and extras:
        out["extras"] = extras
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        times = time_oracle(BATCH, 6)[1:]
        dt = sum(times) / len(times)
        out["cpu_baseline"] = {"value": BATCH * N_STEPS / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": CPU_PROTOCOL.format(tv=torch.__version__, cores=cores, k=len(times), w=1)}
    if rank == 0:
        emit(out_fd, out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
