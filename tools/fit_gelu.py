"""Offline derivation + check of the one-exponential erf-GELU used in the fused GEGLU epilogue (csrc/common.cuh::gelu_erf_fast):
erf(a) = 1 - 2^(-a P(a)), P of degree 4, weighted-minimax (Lawson) fit of -log2(erfc(a)) on [0, 4.2]; prints the coefficients with the
1/sqrt(2) of GELU folded in and the maximum error of the float32 evaluation against scipy's erf.  CPU only."""
import numpy as np
from scipy.special import erf, erfc

x = np.linspace(0, 4.2, 20001)
P = -np.log2(np.maximum(erfc(x), 1e-300))
w = np.log(2) * 2.0 ** (-P)                     # d erf / d P
V = np.vander(x, 6, increasing=True)[:, 1:]      # no constant term: P(0) = 0
lw = np.ones_like(x)
for _ in range(60):
    c = np.linalg.lstsq(V * (w * lw)[:, None], P * w * lw, rcond=None)[0]
    err = np.abs((1 - 2.0 ** (-(V @ c))) - erf(x))
    lw *= 1 + err / err.max()
    lw /= lw.mean()
d = (c * (1 / np.sqrt(2)) ** np.arange(1, 6)).astype(np.float32)
print("coefficients of a^1..a^5 with a = |x| (GELU argument):", [float(v) for v in d])
g = np.linspace(-12, 12, 2000001).astype(np.float32)
a = np.minimum(np.abs(g), np.float32(8.0))
q = np.zeros_like(a)
for k in range(4, -1, -1):
    q = q * a + d[k]
e = np.exp2(-(q * a)).astype(np.float32)
r = (np.float32(0.5) * g * e).astype(np.float32)
gelu = np.where(g >= 0, g - r, r)
ref = 0.5 * g.astype(np.float64) * (1 + erf(g.astype(np.float64) / np.sqrt(2)))
print("max |gelu - exact| =", float(np.abs(gelu - ref).max()))
