"""The erfc fit behind gelu_fast32 (beso_b200/csrc/fast_forward.cu): erfc(a) ~ t q(t) exp(-a^2), t = 1 / (1 + p a).

    python tools/fit_gelu.py

Prints the coefficients (weighted least squares of the absolute erfc error on [0, 7]) and the error of the resulting GELU
evaluated in emulated fp32 against the exact function, next to the error of 0.5 x (1 + erf(x / sqrt 2)) in fp32."""
import numpy as np
from scipy.special import erf, erfc, erfcx

P = 0.3275911


def main():
    a = np.linspace(0, 7, 400001)
    t = 1 / (1 + P * a)
    w = np.exp(-a * a)
    deg = 7
    A = np.stack([t ** k for k in range(1, deg + 1)], 1) * w[:, None]
    c, *_ = np.linalg.lstsq(A, erfcx(a) * w, rcond=None)
    print("coefficients of t^1..t^7:", [float("%.9g" % v) for v in c])
    print("max |erfc error| (float64 evaluation): %.2e" % np.abs(A @ c - erfc(a)).max())
    f32 = np.float32
    x = np.linspace(-10, 10, 2000001)
    a32 = (np.abs(x).astype(f32) * f32(0.70710678118654752440)).astype(f32)
    t32 = (f32(1) / (f32(P) * a32 + f32(1))).astype(f32)
    q = np.full_like(t32, f32(c[6]))
    for k in range(5, -1, -1):
        q = (q * t32 + f32(c[k])).astype(f32)
    e = np.exp2((a32 * a32 * f32(-1.4426950408889634)).astype(f32)).astype(f32)
    half = (f32(0.5) * (q * t32).astype(f32) * e).astype(f32)
    g = (x.astype(f32) * np.where(x >= 0, f32(1) - half, half)).astype(f32)
    exact = x * 0.5 * erfc(-x / np.sqrt(2))
    ref32 = (f32(0.5) * x.astype(f32) * (f32(1) + erf((x.astype(f32) * f32(0.70710678118654752440)).astype(np.float64)).astype(f32))).astype(f32)
    print("max |gelu_fast32 - exact| over [-10, 10]: %.2e" % np.abs(g - exact).max())
    print("max |fp32 0.5 x (1 + erf) - exact|      : %.2e" % np.abs(ref32 - exact).max())


if __name__ == "__main__":
    main()
