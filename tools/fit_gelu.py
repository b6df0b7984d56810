"""Fits the single-MUFU GELU used by the GEMM epilogue (csrc/gemm_tc.cu: gelu_fast).

gelu(v) = max(v, 0) - 0.5*|v|*erfc(|v|/sqrt(2));  erfc(a/sqrt(2)) ~= 2^(-Q(a)),  Q(a) = a*(c1 + a*(c2 + ... )).
Prints the coefficients and the max abs error of the GELU value against scipy's erfc (fp64 and fp32 Horner)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erfc

a = np.linspace(0, 9, 20001)
exact = 0.5 * a * erfc(a / np.sqrt(2))
target = -np.log2(np.maximum(erfc(a / np.sqrt(2)), 1e-300))
deg = 3   # 5 -> 6e-7, 4 -> 9e-6 (not monotone beyond |v| ~ 11), 3 -> 9e-5 with all coefficients positive (shipped)
m = a <= 6.5
V = np.stack([a[m] ** k for k in range(1, deg + 1)], 1)
w = (a[m] * np.exp2(-target[m]))[:, None]
c = np.linalg.lstsq(V * w, target[m] * w[:, 0], rcond=None)[0]


def resid(c):
    q = sum(c[k - 1] * a ** k for k in range(1, deg + 1))
    return 0.5 * a * np.exp2(-np.clip(q, -50, 200)) - exact


for p in (2, 4, 8, 16):  # approach the minimax fit through increasing p-norms
    c = least_squares(lambda c: np.sign(resid(c)) * np.abs(resid(c)) ** (p / 2), c, method='lm', max_nfev=4000).x
af = a.astype(np.float32)
q = np.zeros_like(af)
for k in range(deg, 0, -1):
    q = af * (np.float32(c[k - 1]) + q)
print(f'coefficients c1..c{deg}:', [float('%.9g' % x) for x in c])
print('max |gelu error| fp64 %.3g, fp32 Horner %.3g' % (np.abs(resid(c)).max(), np.abs(np.float32(0.5) * af * np.exp2(-q) - exact).max()))
