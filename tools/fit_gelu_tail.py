"""Fit of q(a) = log2(1 - Phi(a)) used by geglu_pair (csrc/common.cuh): Chebyshev least squares with Lawson
re-weighting towards the minimax ABSOLUTE error of a * 2^q(a), then the fp32 Horner evaluation is checked.
  python tools/fit_gelu_tail.py [degree]"""
import sys

import numpy as np
from numpy.polynomial import chebyshev as Ch
from numpy.polynomial import polynomial as P
from scipy.special import erfc, log_ndtr

AM = 6.0
deg = int(sys.argv[1]) if len(sys.argv) > 1 else 6
a = np.linspace(0, AM, 20001)
q = log_ndtr(-a) / np.log(2)
tail = 0.5 * erfc(a / np.sqrt(2))
w = np.maximum(a, 1e-3) * tail * np.log(2)
w /= w.max()
lw = np.ones_like(a)
t = 2 * a / AM - 1
for _ in range(60):
    c = Ch.chebfit(t, q, deg, w=w * lw)
    err = a * np.exp2(Ch.chebval(t, c)) - a * tail
    e = np.abs(err)
    lw = lw * (1 + 2 * e / e.max())
    lw /= lw.mean()
mono, acc = np.zeros(1), np.array([1.0])
for ck in Ch.cheb2poly(c):
    mono = P.polyadd(mono, ck * acc)
    acc = P.polymul(acc, np.array([-1.0, 2 / AM]))
a32 = a.astype(np.float32)
r = np.float32(mono[-1]) * np.ones_like(a32)
for ck in mono[-2::-1]:
    r = r * a32 + np.float32(ck)
print("degree", deg, "max abs err (fp64 eval)", np.abs(err).max(), "(fp32 Horner)", np.abs((a32 * np.exp2(r)).astype(np.float64) - a * tail).max())
print("coefficients c0..cN:", [float(np.float32(x)) for x in mono])
