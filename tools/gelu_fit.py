"""Coefficients of the GELU used by the GEMM epilogue (csrc/gemm_tcgen05.cu: gelu_erf, gelu_erf2).

gelu(x) = relu(x) - |x| * Phi(-|x|),  Phi(-u) = 2^(-u*Q(u) - 1),  Q = polynomial fit of -log2(2*Phi(-u))/u on (0, 6].
The fit is re-weighted (Lawson iterations) until the error of gelu relative to max(|gelu|, 1e-3) is equal-ripple.
Prints the coefficients (highest power first) in u and in na = -u, and the error of an fp32 evaluation of the kernel's sequence.
usage: python tools/gelu_fit.py [degree=4]"""
import sys

import numpy as np
from scipy.special import erf, log_ndtr, ndtr

d = int(sys.argv[1]) if len(sys.argv) > 1 else 4
u = np.linspace(0, 6, 40001)[1:]
q = -(log_ndtr(-u) / np.log(2) + 1.0) / u
w = np.ones_like(u)
best = None
for _ in range(200):
    c = np.polyfit(u, q, d, w=w)
    t, tt = 2.0 ** (-u * np.polyval(c, u) - 1.0), ndtr(-u)
    err = np.abs(u * (t - tt)) / np.maximum(np.abs(u * tt), 1e-3)
    if best is None or err.max() < best[0]:
        best = (err.max(), c.copy())
    w = w * (1 + err / err.max()) ** 0.5
c = best[1]
cna = c * ((-1.0) ** np.arange(d, -1, -1))
print("degree", d, " fp64 max error relative to max(|gelu|, 1e-3):", best[0])
print("Q(u),  highest power first:", [float(np.float32(v)) for v in c])
print("Q(na), highest power first:", [float(np.float32(v)) for v in cna])
x = np.linspace(-9, 9, 2000001).astype(np.float32)
na = np.maximum(-np.abs(x), np.float32(-6.0))
qq = np.full_like(na, np.float32(cna[0]))
for v in cna[1:]:
    qq = (qq * na + np.float32(v)).astype(np.float32)
arg = (na * qq - np.float32(1.0)).astype(np.float32)
gel = (na * np.exp2(arg.astype(np.float64)).astype(np.float32) + np.maximum(x, 0)).astype(np.float32)
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("fp32 evaluation: max relative (floor 1e-3)", (np.abs(gel - ref) / np.maximum(np.abs(ref), 1e-3)).max(), " max absolute", np.abs(gel - ref).max())
