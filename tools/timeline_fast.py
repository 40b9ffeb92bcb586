"""GPU diagnostics: where one CTA of the FAST kernel spends its second model evaluation.

    python tools/timeline_fast.py [T16|K256|KITCHEN|PUSH] [batch] [fast|precise|stacked] > gpurun_out/timeline.txt

(precise = the precise mode with the 128-row tile layout forced where the shape has it, stacked = with the stacked 64-row
layout forced.)

Prints, in SM clock cycles, (a) the compute warps' phase durations and waits, (b) for the MMA
issuer, per GEMM job, time spent waiting on barriers (compute) vs on the weight ring (producer).
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beso_b200 import BLOCKPUSH_CKPT, K256, KITCHEN_CKPT, T16, _lib   # noqa: E402
from beso_b200.denoiser import build_denoiser                 # noqa: E402
from beso_b200.sampling import get_sigmas_exponential, sample_ddim  # noqa: E402
from beso_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "K256"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    cfg = {"K256": K256, "T16": T16, "KITCHEN": KITCHEN_CKPT, "PUSH": BLOCKPUSH_CKPT}[name]
    dev = torch.device("cuda:0")
    mode = sys.argv[3] if len(sys.argv) > 3 else "fast"
    stacked = mode == "stacked" or cfg.d > 256
    if mode in ("precise", "stacked"):
        os.environ["BESO_PREC_LAYOUT"] = "stacked" if stacked else "p128"      # read by the library at its first launch
        mode = "precise"
    m = build_denoiser(cfg, dev, mode=mode, state_dict=synthetic_state_dict(cfg, 1))
    x = {k: v.to(dev) for k, v in synthetic_inputs(cfg, B, seed=2).items()}
    sig = get_sigmas_exponential(4, 0.005, 1.0)
    L = cfg.n_layers
    # single-accumulator schedules ([Q|K] + V jobs): the 384-column geometry and the precise mode on full tiles (P128)
    wide = cfg.d > 256 or (mode == "precise" and not stacked)
    hsp = 32 if cfg.d // cfg.n_heads <= 32 else 64
    npass = -(-cfg.n_heads // (64 // hsp))
    nch = 12 if cfg.d > 256 else 8
    NF = 8 + 64 * L if wide else 4 + 104 * L + 4
    tl = torch.zeros(6 * NF + 4096, dtype=torch.int64, device=dev)
    sample_ddim(m, x["state"], x["noise"], x["goal"], sig)            # warm-up
    _lib.lib().beso_debug_set_timeline(C.c_void_p(tl.data_ptr()))
    sample_ddim(m, x["state"], x["noise"], x["goal"], sig)
    torch.cuda.synchronize()
    _lib.lib().beso_debug_set_timeline(None)
    tl = tl.cpu().tolist()
    NJ = 2 + L * ((3 if wide else 2) * npass + 2 * nch)
    jobs = [tl[4 * j:4 * j + 4] for j in range(NJ)]          # per MMA job: start, barrier-wait, ring-wait, end
    ev = [v for v in tl[6 * NF:] if v]
    t0 = min(ev[0], jobs[0][0])
    print(f"# {name} B={B} mode={mode}: one evaluation = {max(ev[-1], jobs[-1][3]) - t0} cycles")

    # ---- compute warps ----
    it = iter(ev)
    def nxt():
        return next(it) - t0
    print("## compute warp 0 (cycles since evaluation start)")
    a, b = nxt(), nxt()
    print(f"embed build: {b - a}")
    last = b
    tot = dict(wait=0, ln=0, drain=0, attn=0, sync=0, gelu=0)
    for l in range(L):
        s, e = nxt(), nxt()
        print(f"L{l} LN1: wait {s - last:6d}  run {e - s:6d}")
        tot["wait"] += s - last; tot["ln"] += e - s
        last = e
        for h in range(npass):
            d0, a0, a1, a2 = nxt(), nxt(), nxt(), nxt()
            print(f"L{l} head{h}: wait_acc {d0 - last:6d}  drain+sync {a0 - d0:6d}  attention {a1 - a0:6d}  sync {a2 - a1:6d}")
            tot["wait"] += d0 - last; tot["drain"] += a0 - d0; tot["attn"] += a1 - a0; tot["sync"] += a2 - a1
            last = a2
        s, e = nxt(), nxt()
        print(f"L{l} LN2: wait {s - last:6d}  run {e - s:6d}")
        tot["wait"] += s - last; tot["ln"] += e - s
        last = e
        row = []
        for ch in range(nch):
            s, e = nxt(), nxt()
            row.append(f"w{s - last}/g{e - s}")
            tot["wait"] += s - last; tot["gelu"] += e - s
            last = e
        print(f"L{l} FC1 chunks (wait/gelu): " + " ".join(row))
    s, e, f = nxt(), nxt(), nxt()
    print(f"ln_f: wait {s - last} run {e - s}; head wait {f - e}")
    print("compute totals:", tot)

    # ---- MMA issuer ----
    print("## MMA issuer per job: start, cycles waiting on compute barriers / on the weight ring, total issue time, gap to next job")
    names = ["EMB"]
    layer = []
    for h in range(npass + 1):
        if h < npass:
            layer.append(f"QK{h}" if wide else f"QKV{h}")
        if wide and h < npass:
            layer.append(f"V{h}")
        if h >= 1:
            layer.append(f"PROJ{h - 1}")
    for c in range(nch + 1):
        if c < nch:
            layer.append(f"FC1_{c}")
        if c >= 1:
            layer.append(f"FC2_{c - 1}")
    for l in range(L):
        names += [f"L{l}.{n}" for n in layer]
    names += ["HEAD"]
    tb = tr = 0
    for j, n in enumerate(names):
        st, bw, rw, en = jobs[j]
        tb += bw; tr += rw
        gap = (jobs[j + 1][0] - en) if j + 1 < NJ else 0
        if n.startswith("L0.") or n.startswith("L1.") or not n.startswith("L"):
            print(f"{n:10s} start {st - t0:7d}  barrier {bw:6d}  ring {rw:6d}  total {en - st:6d}  busy {en - st - bw - rw:6d}  gap {gap:5d}")
    print(f"MMA issuer totals: barrier-wait {tb}  ring-wait {tr}  of {jobs[-1][3] - jobs[0][0]}")


if __name__ == "__main__":
    main()
