import sys, time, torch
sys.path.insert(0, ".")
from beso_b200 import K256, sampling
from beso_b200.denoiser import build_denoiser
from beso_b200.synth import synthetic_inputs, synthetic_state_dict
dev = torch.device("cuda:0")
for mode in ("fast", "precise"):
    m = build_denoiser(K256, dev, mode=mode, state_dict=synthetic_state_dict(K256, 1))
    x = {k: v.to(dev) for k, v in synthetic_inputs(K256, 1, seed=2).items()}
    sig = sampling.get_sigmas_exponential(10, 0.005, 1.0)
    with torch.no_grad():
        for _ in range(20): m(x["state"], x["action"], x["goal"], x["sigma"])
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(500): m(x["state"], x["action"], x["goal"], x["sigma"])
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(mode, "forward b1: host issue %.1f us/call, wall %.1f us/call" % ((t1 - t0) / 500 * 1e6, (t2 - t0) / 500 * 1e6))
        for _ in range(5): sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(100): sampling.sample_ddim(m, x["state"], x["noise"], x["goal"], sig)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(mode, "ddim10 b1: host issue %.1f us/call, wall %.1f us/call" % ((t1 - t0) / 100 * 1e6, (t2 - t0) / 100 * 1e6))
