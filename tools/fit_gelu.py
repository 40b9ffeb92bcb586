"""The fit behind gelu_fast32 (beso_b200/csrc/fast_forward.cu): erfc(a) ~ 2^p(t), t = a / 2.85 - 1, a in [0, 5.7].

    python tools/fit_gelu.py

Prints the degree-8 coefficients of p in powers of t (weighted least squares of log2 erfc on a Chebyshev basis, weight =
erfc, i.e. the ABSOLUTE error of erfc is what is minimised) and the error of the resulting GELU evaluated in emulated
fp32 against the exact function, next to the error of 0.5 x (1 + erf(x / sqrt 2)) in fp32."""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

A = 5.7


def main():
    a = np.linspace(0, A, 400001)
    t = 2 * a / A - 1
    c = C.chebfit(t, np.log2(erfc(a)), 8, w=np.maximum(erfc(a), 1e-12))
    pc = C.cheb2poly(c)
    print("coefficients of t^0..t^8:", [float("%.9g" % v) for v in pc])
    print("max |erfc error| (float64 evaluation): %.2e" % np.abs(np.exp2(C.chebval(t, c)) - erfc(a)).max())
    f32 = np.float32
    x = np.linspace(-10, 10, 2000001)
    a32 = np.minimum((np.abs(x).astype(f32) * f32(0.70710678118654752440)).astype(f32), f32(A))
    t32 = (a32 * f32(2 / A) - f32(1)).astype(f32)
    p = np.full_like(t32, f32(pc[8]))
    for k in range(7, -1, -1):
        p = (p * t32 + f32(pc[k])).astype(f32)
    half = (f32(0.5) * np.exp2(p.astype(np.float64)).astype(f32)).astype(f32)
    g = (x.astype(f32) * np.where(x >= 0, f32(1) - half, half)).astype(f32)
    exact = x * 0.5 * erfc(-x / np.sqrt(2))
    ref32 = (f32(0.5) * x.astype(f32) * (f32(1) + erf((x.astype(f32) * f32(0.70710678118654752440)).astype(np.float64)).astype(f32))).astype(f32)
    print("max |gelu_fast32 - exact| over [-10, 10]: %.2e" % np.abs(g - exact).max())
    print("max |fp32 0.5 x (1 + erf) - exact|      : %.2e" % np.abs(ref32 - exact).max())


if __name__ == "__main__":
    main()
