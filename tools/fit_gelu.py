"""Fit used by geglu2() in csrc/idb_common.cuh: Phi(g) = 0.5 (1 + erf(g / sqrt 2)) ~ 0.5 + g Q((g / 4.25)^2) on
|g| <= 4.25 (clamped outside), Q of degree 8, near-minimax by Lawson-weighted least squares; prints the
monomial coefficients (c0 .. c8) and the fp32 Horner error."""
import numpy as np
from scipy.special import erf

GMAX, DEG = 4.25, 8
Phi = lambda g: 0.5 * (1 + erf(g / np.sqrt(2)))
n = 6000
g = np.linspace(1e-6, GMAX, n)
s = (g / GMAX) ** 2
target = (Phi(g) - 0.5) / g
T = np.polynomial.chebyshev.chebvander(2 * s - 1, DEG)
w = np.ones(n)
for _ in range(80):
    A = T * np.sqrt(w)[:, None] * g[:, None]
    c, *_ = np.linalg.lstsq(A, target * np.sqrt(w) * g, rcond=None)
    e = np.abs((T @ c - target) * g)
    w = w * (e / e.max() + 1e-3)
    w /= w.sum()
mono = np.polynomial.polynomial.polyfit(s, T @ c, DEG)
gg = np.linspace(-8, 8, 200001).astype(np.float32)
gc = np.clip(gg, -GMAX, GMAX).astype(np.float32)
ss = ((gc * np.float32(1 / GMAX)) ** 2).astype(np.float32)
m32 = mono.astype(np.float32)
acc = np.full_like(ss, m32[-1])
for k in range(DEG - 1, -1, -1):
    acc = (acc * ss + m32[k]).astype(np.float32)
ph = (gc * acc + np.float32(0.5)).astype(np.float32)
print("max |Phi error| (fp32 Horner):", np.abs(ph.astype(np.float64) - Phi(gg.astype(np.float64))).max())
print(", ".join(f"{v:.9e}f" for v in mono))
